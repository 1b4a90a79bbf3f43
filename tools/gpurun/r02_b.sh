#!/bin/bash
# round 2, second GPU pass: the dense tensor-core path (tests first, then bench lines and a launch list)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "dense or switches or default_visual or published_bench" 2>&1 | tail -15
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for v in max 10.0; do
  timeout 400 python bench.py --visual-threshold $v --no-cpu-baseline > gpurun_out/r02b_bench_cfg5_thr$v.json 2> gpurun_out/r02b_bench_cfg5_thr$v.err
done
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02b_bench_cfg5.json 2> gpurun_out/r02b_bench_cfg5.err
timeout 300 python bench.py --config cfg2 --no-cpu-baseline > gpurun_out/r02b_bench_cfg2.json 2> gpurun_out/r02b_bench_cfg2.err
python - <<'PY'
import json
for c in ("cfg5_thrmax", "cfg5_thr10.0", "cfg5", "cfg2"):
    try:
        d = json.loads(open(f"gpurun_out/r02b_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, "%.4e" % d["value"], "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 3),
              "launches/step", d.get("gpu_launches_per_step"), {k: round(v, 4) for k, v in d.get("stages_ms", {}).items()},
              "frac", d.get("roofline", {}).get("frac"), d.get("clocks"))
    except Exception as e:
        print(c, "failed", e)
PY
tail -3 gpurun_out/r02b_bench_cfg5_thrmax.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches_cfg5_thrmax.csv \
  python bench.py --steps 2 --warmup 4 --no-cpu-baseline --visual-threshold max > gpurun_out/r02b_ncu_bench.log 2>&1
tail -1 gpurun_out/r02b_ncu_bench.log | cut -c1-200
