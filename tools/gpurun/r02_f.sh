#!/bin/bash
# round 2, sixth GPU pass: suite, dense bench lines, steady-state full captures of one frame per config
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -rf > gpurun_out/r02f_pytest.log 2>&1
tail -3 gpurun_out/r02f_pytest.log
grep -E "^(FAILED|ERROR)" gpurun_out/r02f_pytest.log | cut -c1-200 | head -20
for v in max 10.0; do
  timeout 300 python bench.py --visual-threshold $v --no-cpu-baseline > gpurun_out/r02f_bench_thr$v.json 2> gpurun_out/r02f_bench_thr$v.err
done
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02f_bench_cfg5.json 2> gpurun_out/r02f_bench_cfg5.err
python - <<'PY'
import json
for c in ("thrmax", "thr10.0", "cfg5"):
    try:
        d = json.loads(open(f"gpurun_out/r02f_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, "%.4e" % d["value"], "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 3),
              {k: round(v, 4) for k, v in d.get("stages_ms", {}).items()}, "frac", d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(c, "failed", e)
PY
# full captures of one steady-state frame (every kernel of the frame, nothing else)
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o gpurun_out/r02f_frame_cfg5 \
  python tools/profile_frame.py cfg5 9 > gpurun_out/r02f_ncu_cfg5.log 2>&1
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o gpurun_out/r02f_frame_cfg5_dense \
  python tools/profile_frame.py cfg5 9 --visual-threshold max > gpurun_out/r02f_ncu_cfg5_dense.log 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o gpurun_out/r02f_frame_cfg4 \
  python tools/profile_frame.py cfg4 9 > gpurun_out/r02f_ncu_cfg4.log 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o gpurun_out/r02f_frame_cfg2 \
  python tools/profile_frame.py cfg2 9 > gpurun_out/r02f_ncu_cfg2.log 2>&1
ls -la gpurun_out/r02f_frame_*.ncu-rep
tail -2 gpurun_out/r02f_ncu_cfg5.log | cut -c1-300
