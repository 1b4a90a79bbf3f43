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
for cfg in 5 3; do
  timeout 300 python bench.py --no-cpu-baseline --config cfg$cfg --steps 20 --warmup 5 > gpurun_out/r02n_cfg$cfg.json 2> /dev/null
  show gpurun_out/r02n_cfg$cfg.json cfg$cfg
done
timeout 600 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -2
