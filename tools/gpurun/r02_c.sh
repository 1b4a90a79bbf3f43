#!/bin/bash
# round 2, third GPU pass: dense path after the epilogue / sample / select rewrites, cp.async refine, lazy dense positional
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
SB200_TRACE=1 timeout 400 python bench.py --visual-threshold max --no-cpu-baseline > gpurun_out/r02c_bench_cfg5_thrmax.json 2> gpurun_out/r02c_bench_cfg5_thrmax.err
grep "absorbed" gpurun_out/r02c_bench_cfg5_thrmax.err | tail -12
grep "predict:" gpurun_out/r02c_bench_cfg5_thrmax.err | tail -6
timeout 400 python bench.py --visual-threshold 10.0 --no-cpu-baseline > gpurun_out/r02c_bench_cfg5_thr10.0.json 2> gpurun_out/r02c_bench_cfg5_thr10.0.err
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02c_bench_cfg5.json 2> gpurun_out/r02c_bench_cfg5.err
SB200_REFINE=regs timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02c_bench_cfg5_refregs.json 2> gpurun_out/r02c_bench_cfg5_refregs.err
for c in cfg2 cfg3 cfg4; do
  timeout 300 python bench.py --config $c --no-cpu-baseline > gpurun_out/r02c_bench_$c.json 2> gpurun_out/r02c_bench_$c.err
done
python - <<'PY'
import json
for c in ("cfg5_thrmax", "cfg5_thr10.0", "cfg5", "cfg5_refregs", "cfg2", "cfg3", "cfg4"):
    try:
        d = json.loads(open(f"gpurun_out/r02c_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, "%.4e" % d["value"], "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 3),
              "launches/step", d.get("gpu_launches_per_step"), {k: round(v, 4) for k, v in d.get("stages_ms", {}).items()},
              "frac", d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(c, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02c_launches_cfg5_thrmax.csv \
  python bench.py --steps 2 --warmup 6 --no-cpu-baseline --visual-threshold max > gpurun_out/r02c_ncu_bench.log 2>&1
tail -1 gpurun_out/r02c_ncu_bench.log | cut -c1-200
