#!/bin/bash
mkdir -p gpurun_out
show() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.4e" % d["value"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0) or 0, 3),
          "host", {k: round(v, 3) for k, v in d.get("host_ms_per_step", {}).items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for rep in 1 2 3; do
  timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --visual-threshold max > gpurun_out/r02r_thrmax_$rep.json 2>/dev/null
  show gpurun_out/r02r_thrmax_$rep.json "thr max rep $rep"
done
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02r_cfg5.json 2>/dev/null
show gpurun_out/r02r_cfg5.json "cfg5"
timeout 600 python -m pytest tests/test_gpu_tracker.py -m gpu -x -q 2>&1 | tail -2
