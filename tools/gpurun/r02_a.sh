#!/bin/bash
# round 2, first GPU pass: parity suite on the stream-ordered engine, smoke, bench lines (cfg5 default + cfg2/3/4),
# ncu launch list of a short bench run
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
nproc
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r02a_bench_cfg5.json 2> gpurun_out/r02a_bench_cfg5.err
for c in cfg2 cfg3 cfg4; do
  timeout 300 python bench.py --config $c --no-cpu-baseline > gpurun_out/r02a_bench_$c.json 2> gpurun_out/r02a_bench_$c.err
done
python - <<'PY'
import json
for c in ("cfg5", "cfg2", "cfg3", "cfg4"):
    try:
        d = json.loads(open(f"gpurun_out/r02a_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, "%.4e" % d["value"], "ms/step", round(d["ms_per_step"], 4), "e2e %.4e" % d["e2e"]["value"], round(d["e2e"]["ms_per_step"], 3),
              "launches/step", d.get("gpu_launches_per_step"), {k: round(v, 4) for k, v in d.get("stages_ms", {}).items()},
              "frac", d.get("roofline", {}).get("frac"), d.get("clocks"), d.get("cpu_baseline"))
    except Exception as e:
        print(c, "failed", e)
PY
tail -3 gpurun_out/r02a_bench_cfg5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_cfg5.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_ncu_bench.log 2>&1
tail -2 gpurun_out/r02a_ncu_bench.log | cut -c1-300
