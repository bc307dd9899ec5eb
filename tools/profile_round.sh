#!/bin/bash
# One profiling pass of the default bench command (run under gpurun, one GPU):
#   1. the plain bench line (never taken under ncu)
#   2. ncu launch list (gpu__time_duration per launch) of the same command
#   3. ncu --set full of one launch of every kernel of the step (DRAM traffic, pipe utilisation, source page)
set -x
tag=${1:-v}
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || exit 1
tail -c 600 gpurun_out/bench_$tag.json
DH_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1
# device-resident launches of 512 frames come first (warm-up steps): skip the first pass, take one launch of each kernel
ncu --set full --import-source on --clock-control none -k regex:"box_image|traverse_kernel|gate_coarse|box_build|meanshift_kernel|seed_kernel" -s 12 -c 6 \
    -o gpurun_out/prof_step_$tag -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"biwi_decode" -s 1 -c 1 \
    -o gpurun_out/prof_biwi_$tag -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_biwi_$tag.log 2>&1
ls -la gpurun_out/*$tag*
