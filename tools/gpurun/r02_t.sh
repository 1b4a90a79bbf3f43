#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02t_cfg5.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02t_cfg5.json").read().strip().splitlines()[-1])
print("cfg5 value %.4e" % d["value"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3), {k: round(v, 3) for k, v in d["stages_ms"].items()}, "frac", round(d["roofline"]["frac"], 3))
PY
