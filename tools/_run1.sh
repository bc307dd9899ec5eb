set -x
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "box_image or variants" > gpurun_out/t_box.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_box.log
tail -3 gpurun_out/t_box.log
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/bench_f0.json 2> gpurun_out/bench_f0.err
python tools/benchline.py gpurun_out/bench_f*.json
