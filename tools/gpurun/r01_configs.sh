#!/bin/bash
# per-config bench lines (cfg2..cfg5), host-side trace of predict, NMS 10k timing
mkdir -p gpurun_out
for c in cfg2 cfg3 cfg4; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  tail -c 600 gpurun_out/bench_$c.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]) if False else None
" 2>/dev/null
done
SB200_TRACE=1 timeout 300 python bench.py --steps 6 --warmup 4 --no-cpu-baseline > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err
timeout 300 python tools/nms_bench.py 10 --check > gpurun_out/nms_10k.json 2> gpurun_out/nms_10k.err
timeout 400 python bench.py --steps 40 --warmup 4 --no-cpu-baseline > gpurun_out/bench_k40.json 2> gpurun_out/bench_k40.err
python - <<'PY'
import json
for c in ("cfg2", "cfg3", "cfg4", "trace", "k40"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{c}.json").read().strip().splitlines()[-1])
        print(c, "%.3e" % d["value"], round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4),
              {k: round(v, 4) for k, v in d["stages_ms"].items()}, round(d["roofline"]["frac"], 4), d["roofline"]["unit"])
    except Exception as e:
        print(c, "failed", e)
print(open("gpurun_out/nms_10k.json").read())
PY
grep "predict: setup" gpurun_out/bench_trace.err | tail -8
tail -2 gpurun_out/nms_10k.err
