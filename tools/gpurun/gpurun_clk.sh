#!/bin/bash
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 50 > gpurun_out/clk.log 2>&1 &
SMI=$!
timeout 600 python bench.py --steps 100 --warmup 4 --no-cpu-baseline > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err
kill $SMI
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_long.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["stages_ms"], d["roofline"]["frac"], d["clocks"])
PY
sort gpurun_out/clk.log | uniq -c | sort -rn | head -12
