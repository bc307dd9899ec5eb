set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_box.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_box.log
tail -5 gpurun_out/t_box.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_box4.json 2> gpurun_out/bench_box4.err
python - <<'PY'
import json
for n in ("box4",):
    try:
        d=json.loads(open("gpurun_out/bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["value"]), d["ms_per_step"], d["stage_ms_per_step"], d["e2e"]["value"])
    except Exception as e: print(n, "ERR", e)
PY
