#!/bin/bash
# round 2, fifth GPU pass: full suite with a complete failure report, templated weight-sum epilogue, bench lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -rf > gpurun_out/r02e_pytest.log 2>&1
tail -5 gpurun_out/r02e_pytest.log
grep -E "^(FAILED|ERROR)" gpurun_out/r02e_pytest.log | cut -c1-200 | head -40
grep -E "^E  " gpurun_out/r02e_pytest.log | sort | uniq -c | sort -rn | head -20 | cut -c1-260
for v in max; do
  for dbg in 0 4 7; do
    SB200_DENSE_DBG=$dbg timeout 300 python bench.py --visual-threshold $v --no-cpu-baseline --steps 6 > gpurun_out/r02e_bench_thr${v}_dbg$dbg.json 2> gpurun_out/r02e_bench_thr${v}_dbg$dbg.err
  done
done
SB200_TRACE=1 timeout 300 python bench.py --visual-threshold max --no-cpu-baseline --steps 4 2>&1 >/dev/null | grep "dense frame" | tail -3
timeout 300 python bench.py --visual-threshold 10.0 --no-cpu-baseline > gpurun_out/r02e_bench_thr10.json 2> gpurun_out/r02e_bench_thr10.err
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02e_bench_cfg5.json 2> gpurun_out/r02e_bench_cfg5.err
for c in cfg2 cfg3 cfg4; do
  timeout 300 python bench.py --config $c --no-cpu-baseline > gpurun_out/r02e_bench_$c.json 2> gpurun_out/r02e_bench_$c.err
done
python - <<'PY'
import json
for c in ("thrmax_dbg0", "thrmax_dbg4", "thrmax_dbg7", "thr10", "cfg5", "cfg2", "cfg3", "cfg4"):
    try:
        d = json.loads(open(f"gpurun_out/r02e_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, "%.4e" % d["value"], "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 3),
              {k: round(v, 4) for k, v in d.get("stages_ms", {}).items()}, "frac", d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(c, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02e_launches_cfg5_thrmax.csv \
  python bench.py --steps 2 --warmup 7 --no-cpu-baseline --visual-threshold max > gpurun_out/r02e_ncu_bench.log 2>&1
