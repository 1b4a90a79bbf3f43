#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tracker.py tests/test_gpu_api.py tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python bench.py --steps 8 --warmup 4 > gpurun_out/bench_cluster.json 2> gpurun_out/bench_cluster.err; python - <<'PY'
import json
for n in ("cluster",):
    d=json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["stages_ms"], d["roofline"]["frac"])
PY
SB200_SCREEN_SINGLE=1 timeout 300 python bench.py --steps 8 --warmup 4 > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; python - <<'PY'
import json
for n in ("single",):
    d=json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["stages_ms"], d["roofline"]["frac"])
PY
