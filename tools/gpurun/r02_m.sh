#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches_cfg5.csv python bench.py --no-cpu-baseline --steps 2 --warmup 6 > gpurun_out/r02m_ncu_bench.log 2>&1
python - <<'PY'
import csv
rows = []
with open("gpurun_out/r02m_launches_cfg5.csv") as fh:
    lines = [l for l in fh if not l.startswith("==")]
r = csv.DictReader(lines)
for x in r:
    rows.append((x["Kernel Name"].split("(")[0][:60], float(x["Metric Value"].replace(",", "")), x["Metric Unit"]))
# last frame: find last frame_setup
idx = [i for i, x in enumerate(rows) if "frame_setup" in x[0]]
a = idx[-2]; b = idx[-1]
tot = 0
for n, v, u in rows[a:b]:
    v = v / 1000 if u.startswith("n") else v
    tot += v
    print(f"{n:62s} {v:8.2f} us")
print("sum", round(tot, 1), "us over", b - a, "launches")
PY
for t in 1 2 3; do
timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 5 --visual-threshold max > gpurun_out/r02m_thrmax.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02m_thrmax.json").read().strip().splitlines()[-1])
print("thrmax ms/step", round(d["ms_per_step"], 4), {k: round(v, 3) for k, v in d["stages_ms"].items()}, {k: round(v, 3) for k, v in d.get("host_ms_per_step", {}).items()})
PY
done
