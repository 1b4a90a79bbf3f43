#!/bin/bash
# round 2: apply split + side-stream overlap: parity, A/B bench, initcheck re-run
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02i_pytest.log
for side in 1 0; do
  SB200_SIDE_STREAM=$side timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02i_cfg5_side$side.json 2> gpurun_out/r02i_cfg5_side$side.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02i_cfg5_side$side.json").read().strip().splitlines()[-1])
print("side=$side ms/step", round(d["ms_per_step"], 4), "e2e", d["e2e"].get("ms_per_step"), {k: round(v, 3) for k, v in d["stages_ms"].items()}, "launches", d.get("gpu_launches"))
PY
done
for cfg in 2 3 4; do
  timeout 300 python bench.py --no-cpu-baseline --config cfg$cfg --steps 20 --warmup 5 > gpurun_out/r02i_cfg$cfg.json 2> /dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02i_cfg$cfg.json").read().strip().splitlines()[-1])
print("cfg$cfg ms/step", round(d["ms_per_step"], 4), {k: round(v, 3) for k, v in d["stages_ms"].items()})
PY
done
timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 5 --visual-threshold max > gpurun_out/r02i_thrmax.json 2>/dev/null
python - <<PY
import json
d = json.loads(open("gpurun_out/r02i_thrmax.json").read().strip().splitlines()[-1])
print("thrmax ms/step", round(d["ms_per_step"], 4), {k: round(v, 3) for k, v in d["stages_ms"].items()})
PY
timeout 900 compute-sanitizer --tool initcheck --print-limit 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_own_area.py tests/test_gpu_api.py -m gpu -q -x > gpurun_out/r02i_initcheck.log 2>&1
grep "    at sb::" gpurun_out/r02i_initcheck.log | sed 's/(.*)+0x[0-9a-f]*//' | sort | uniq -c | head -20
grep 'ERROR SUMMARY\|passed\|failed' gpurun_out/r02i_initcheck.log | tail -3
