#!/bin/bash
# One profiling pass of the bench command (run under gpurun, one GPU):
#   1. the plain bench line (never taken under ncu) and the build id of the library
#   2. ncu launch list (gpu__time_duration per launch) of the same command
#   3. the plain run, then ncu --set full, of one 512-frame launch of every kernel of the step
#      (DRAM traffic, pipe utilisation, source page) + a --metrics pass of the same launches for
#      the wavefront counters the full set does not hold
# Decode here with tools/ncu_summary.py (writes profiles/<tag>_kernels.md and profiles/traverse_profile.json).
set -x
tag=${1:-v}
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
python -c "from depthhead_b200 import capi; print(capi.load().dh_build_id().decode())" > gpurun_out/build_id_$tag.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || exit 1
tail -c 400 gpurun_out/bench_$tag.json
SMALL="--no-cpu-baseline --no-strong --no-extra-forests"
DH_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 $SMALL > gpurun_out/ncu_launches_$tag.log 2>&1
K='regex:box_image|traverse_kernel|patch_gate|gate_coarse|box_build|meanshift_kernel|seed_kernel'
python bench.py --frames 512 --chunk 512 --steps 2 --warmup 3 $SMALL > gpurun_out/bench_prof_$tag.json 2> gpurun_out/bench_prof_$tag.err &&
ncu --set full --import-source on --clock-control none -k "$K" -s 7 -c 7 \
    -o gpurun_out/prof_step_$tag -f python bench.py --frames 512 --chunk 512 --steps 1 --warmup 3 $SMALL > gpurun_out/ncu_full_$tag.log 2>&1
M=l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_tex_wavefronts.sum,l1tex__m_xbar2l1tex_read_sectors.sum,l1tex__t_sectors.sum,l1tex__t_sectors_lookup_hit.sum,l1tex__t_set_accesses.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k "$K" -s 7 -c 7 --csv --log-file gpurun_out/prof_wf_$tag.csv \
    python bench.py --frames 512 --chunk 512 --steps 1 --warmup 3 $SMALL > gpurun_out/ncu_wf_$tag.log 2>&1
ls -la gpurun_out/*$tag*
