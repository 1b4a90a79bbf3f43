#!/bin/bash
# e2e arm with and without binding to the GPU's NUMA node
mkdir -p gpurun_out
SB200_BENCH_NUMA=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_numa0.json 2> gpurun_out/bench_numa0.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_numa1.json 2> gpurun_out/bench_numa1.err
python - <<'PY'
import json
for c in ("numa0", "numa1"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{c}.json").read().strip().splitlines()[-1])
        print(c, "%.4e" % d["value"], round(d["ms_per_step"], 4), "e2e %.4e" % d["e2e"]["value"], round(d["e2e"]["ms_per_step"], 4), d["e2e"].get("host_affinity"))
    except Exception as e:
        print(c, "failed", e)
PY
nvidia-smi topo -m 2>/dev/null | head -6
