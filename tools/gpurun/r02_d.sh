#!/bin/bash
# round 2, fourth GPU pass: dense-path diagnostics (fallback reasons, epilogue toggles, ncu source view of vis_wsum_kernel),
# refine with candidate-row reuse, pipeline-aware capacities
mkdir -p gpurun_out
timeout 300 python tools/dense_probe.py 32 2>&1 | grep -v "predict:" | tail -30
for dbg in 0 1 2 4 7; do
  SB200_DENSE_DBG=$dbg timeout 300 python bench.py --visual-threshold max --no-cpu-baseline --steps 5 > gpurun_out/r02d_bench_thrmax_dbg$dbg.json 2> gpurun_out/r02d_bench_thrmax_dbg$dbg.err
done
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02d_bench_cfg5.json 2> gpurun_out/r02d_bench_cfg5.err
for c in cfg2 cfg4; do
  timeout 300 python bench.py --config $c --no-cpu-baseline > gpurun_out/r02d_bench_$c.json 2> gpurun_out/r02d_bench_$c.err
done
python - <<'PY'
import json
for c in ("thrmax_dbg0", "thrmax_dbg1", "thrmax_dbg2", "thrmax_dbg4", "thrmax_dbg7", "cfg5", "cfg2", "cfg4"):
    try:
        d = json.loads(open(f"gpurun_out/r02d_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, "%.4e" % d["value"], "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 3),
              {k: round(v, 4) for k, v in d.get("stages_ms", {}).items()}, "frac", d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(c, "failed", e)
PY
timeout 900 ncu --set full --import-source on -k regex:vis_wsum --launch-skip 6 -c 1 --clock-control none -f -o gpurun_out/r02d_wsum \
  python bench.py --steps 2 --warmup 6 --no-cpu-baseline --visual-threshold max > gpurun_out/r02d_ncu_wsum.log 2>&1
ls -la gpurun_out/r02d_wsum.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12
