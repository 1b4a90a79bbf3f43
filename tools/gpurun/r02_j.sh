#!/bin/bash
# round 2: host cost per frame (new host counters), side stream A/B, refine claim size A/B
mkdir -p gpurun_out
show() {
python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3), {k: round(v, 3) for k, v in d["stages_ms"].items()},
      "host", {k: round(v, 3) for k, v in d.get("host_ms_per_step", {}).items()})
PY
}
nproc; lscpu | grep -i "model name\|^CPU(s)\|MHz" | head -4; uptime
for rep in 1 2; do
  for side in 1 0; do
    SB200_SIDE_STREAM=$side timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02j_side${side}_$rep.json 2>/dev/null
    show gpurun_out/r02j_side${side}_$rep.json "side=$side rep=$rep"
  done
done
SB200_REFINE_PAIRS=16 timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02j_rp16.json 2>/dev/null
show gpurun_out/r02j_rp16.json "refine pairs 16"
SB200_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 6 --warmup 4 > /dev/null 2> gpurun_out/r02j_trace.err
grep "predict:" gpurun_out/r02j_trace.err | tail -24
