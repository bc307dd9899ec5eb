set -x
ncu --set full --import-source on --clock-control none -k regex:"box_image|traverse_kernel|gate_coarse" -s 9 -c 3 -o gpurun_out/prof_box_v4 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_box_v4.log 2>&1
tail -3 gpurun_out/ncu_box_v4.log
