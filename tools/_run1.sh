set -x
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "box_image or variants" > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_j0.json 2> gpurun_out/bench_j0.err
python tools/benchline.py gpurun_out/bench_j0.json
