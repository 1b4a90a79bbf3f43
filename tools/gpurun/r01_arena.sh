#!/bin/bash
# full GPU parity suite + short and long bench after the feature-arena / end-of-frame sweep / forked positional stage
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
SB200_NO_FORK=1 timeout 600 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -3
for k in 10 40; do
  timeout 400 python bench.py --steps $k --warmup 4 --no-cpu-baseline > gpurun_out/bench_arena_k$k.json 2> gpurun_out/bench_arena_k$k.err
done
SB200_NO_FORK=1 timeout 400 python bench.py --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/bench_arena_nofork.json 2> gpurun_out/bench_arena_nofork.err
timeout 300 python bench.py --config cfg3 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/bench_arena_cfg3.json 2> gpurun_out/bench_arena_cfg3.err
python - <<'PY'
import json
for c in ("k10", "k40", "nofork", "cfg3"):
    try:
        d = json.loads(open(f"gpurun_out/bench_arena_{c}.json").read().strip().splitlines()[-1])
        print(c, "%.3e" % d["value"], round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4),
              {k: round(v, 4) for k, v in d["stages_ms"].items()}, round(d["roofline"]["frac"], 4))
    except Exception as e:
        print(c, "failed", e)
PY
tail -3 gpurun_out/bench_arena_k10.err
