#!/bin/bash
mkdir -p gpurun_out
show() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3), {k: round(v, 3) for k, v in d["stages_ms"].items()},
          "host", {k: round(v, 3) for k, v in d.get("host_ms_per_step", {}).items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02p_pytest.log
for pa in 1 0 1; do
  SB200_PREP_AHEAD=$pa timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02p_cfg5_pa$pa.json 2>gpurun_out/r02p_cfg5_pa$pa.err
  show gpurun_out/r02p_cfg5_pa$pa.json "cfg5 prep_ahead=$pa"
done
for cfg in 2 3 4; do
  timeout 300 python bench.py --no-cpu-baseline --config cfg$cfg --steps 20 --warmup 5 > gpurun_out/r02p_cfg$cfg.json 2> /dev/null
  show gpurun_out/r02p_cfg$cfg.json cfg$cfg
done
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --visual-threshold max > gpurun_out/r02p_thrmax.json 2>/dev/null
show gpurun_out/r02p_thrmax.json thrmax
