#!/bin/bash
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
rc=$?
timeout 600 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_api.py tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -5
for mode in pair multicast; do
SB200_SCREEN=$mode timeout 300 python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$mode.json").read().strip().splitlines()[-1])
print("$mode", d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["stages_ms"], d["roofline"]["frac"])
PY
done
