set -x
timeout 900 python -m pytest tests/test_gpu_biwi.py -m gpu -x -q > gpurun_out/t_biwi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_biwi.log
tail -12 gpurun_out/t_biwi.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m0.json 2> gpurun_out/bench_m0.err
tail -3 gpurun_out/bench_m0.err
python tools/benchline.py gpurun_out/bench_m0.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_m0.json').read().strip().splitlines()[-1]); print(d.get('e2e_biwi'))"
