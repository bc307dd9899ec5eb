"""Print the headline fields of bench JSON lines: python tools/benchline.py gpurun_out/bench_x.json ..."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        st = {k: round(v, 3) for k, v in d.get("stage_ms_per_step", {}).items()}
        print(path, round(d["value"]), round(d["ms_per_step"], 3), st, round(d["e2e"]["value"]))
    except Exception as e:  # noqa: BLE001
        print(path, "ERR", e)
