#!/bin/bash
# round-end check as the driver runs it: GPU parity suite, smoke(), reference arm, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
for c in ("final_ref", "final"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{c}.json").read().strip().splitlines()[-1])
        print(c, "%.4e" % d["value"], round(d["ms_per_step"], 4), "e2e %.4e" % d["e2e"]["value"], d.get("cpu_baseline"),
              {k: round(v, 4) for k, v in d.get("stages_ms", {}).items()}, d.get("roofline", {}).get("frac"), d.get("clocks"))
    except Exception as e:
        print(c, "failed", e)
PY
tail -2 gpurun_out/bench_final.err
