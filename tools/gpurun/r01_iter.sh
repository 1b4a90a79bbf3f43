#!/bin/bash
# GPU parity suite + cfg5 / cfg2 / cfg3 / cfg4 bench lines (iteration check)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for c in cfg5 cfg2 cfg3 cfg4; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/bench_it_$c.json 2> gpurun_out/bench_it_$c.err
done
SB200_NO_FORK=1 timeout 400 python bench.py --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/bench_it_nofork.json 2> gpurun_out/bench_it_nofork.err
SB200_TRACE=1 timeout 300 python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/bench_it_trace.json 2> gpurun_out/bench_it_trace.err
python - <<'PY'
import json
for c in ("cfg5", "nofork", "cfg2", "cfg3", "cfg4"):
    try:
        d = json.loads(open(f"gpurun_out/bench_it_{c}.json").read().strip().splitlines()[-1])
        print(c, "%.3e" % d["value"], round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4),
              {k: round(v, 4) for k, v in d["stages_ms"].items()}, round(d["roofline"]["frac"], 4))
    except Exception as e:
        print(c, "failed", e)
PY
grep "predict: setup" gpurun_out/bench_it_trace.err | tail -4
tail -2 gpurun_out/bench_it_cfg5.err
