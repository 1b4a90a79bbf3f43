#!/bin/bash
mkdir -p gpurun_out
show() {
python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3), {k: round(v, 3) for k, v in d["stages_ms"].items()},
      "host", {k: round(v, 3) for k, v in d.get("host_ms_per_step", {}).items()})
PY
}
SB200_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 5 > /dev/null 2> gpurun_out/r02k_trace.err
grep -v "frame absorbed" gpurun_out/r02k_trace.err | grep "sb200" | tail -30 | cut -c1-170
for rep in 1 2 3; do
  timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02k_cfg5_$rep.json 2>/dev/null
  show gpurun_out/r02k_cfg5_$rep.json "cfg5 rep=$rep"
done
SB200_SIDE_STREAM=0 timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02k_cfg5_side0.json 2>/dev/null
show gpurun_out/r02k_cfg5_side0.json "cfg5 side=0"
for cfg in 2 3 4; do
  timeout 300 python bench.py --no-cpu-baseline --config cfg$cfg --steps 20 --warmup 5 > gpurun_out/r02k_cfg$cfg.json 2> /dev/null
  show gpurun_out/r02k_cfg$cfg.json cfg$cfg
done
timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 5 --visual-threshold max > gpurun_out/r02k_thrmax.json 2>/dev/null
show gpurun_out/r02k_thrmax.json thrmax
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02k_pytest.log
