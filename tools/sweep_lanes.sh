python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_gpu.log
for lanes in 1 2 3 4; do for chunk in 128 256 512; do
DH_LANES=$lanes python bench.py --no-cpu-baseline --forest-from arrays --chunk $chunk > gpurun_out/bench_l${lanes}_c${chunk}.json 2> gpurun_out/bench.err
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_l${lanes}_c${chunk}.json')); print('lanes $lanes chunk $chunk value %.0f e2e %.0f ms %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"
tail -2 gpurun_out/bench.err
done; done
