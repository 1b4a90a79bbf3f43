#!/bin/bash
# round 2: last check of the final build: the whole -m gpu suite, smoke(), bench lines
mkdir -p gpurun_out
show() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.4e" % d["value"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0) or 0, 3),
          {k: round(v, 3) for k, v in d.get("stages_ms", {}).items()}, "frac", round((d.get("roofline") or {}).get("frac", 0) or 0, 3))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02s_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for rep in 1 2; do
  timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02s_cfg5_$rep.json 2>/dev/null
  show gpurun_out/r02s_cfg5_$rep.json "cfg5 rep $rep"
done
for cfg in 2 3 4; do
  timeout 300 python bench.py --no-cpu-baseline --config cfg$cfg --steps 20 --warmup 5 > gpurun_out/r02s_cfg$cfg.json 2> /dev/null
  show gpurun_out/r02s_cfg$cfg.json cfg$cfg
done
