#!/bin/bash
# round 2: compute-sanitizer passes over the small parity tests + a survivor-rate sweep of the headline config
mkdir -p gpurun_out
SEL='tests/test_gpu_parity.py tests/test_gpu_own_area.py tests/test_gpu_api.py'
for tool in memcheck racecheck initcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 5 python -m pytest $SEL -m gpu -q -x \
    > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -c 'Invalid\|Race reported\|Uninitialized\|Barrier error' gpurun_out/r02_sanitizer_$tool.log) finding line(s)"
  grep -m3 -A3 'Invalid\|Race reported\|Uninitialized\|Barrier error' gpurun_out/r02_sanitizer_$tool.log | cut -c1-200
  grep 'ERROR SUMMARY\|passed\|failed' gpurun_out/r02_sanitizer_$tool.log | tail -2
done
# the stream-ordered engine and the dense path under memcheck (one small case each)
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_tracker.py -m gpu -q -x \
  -k "frames_in_flight or (dense_tensor_core and 3-2-0-3.4 and 3]) or ragged or history" > gpurun_out/r02_sanitizer_memcheck_tracker.log 2>&1
grep 'ERROR SUMMARY\|passed\|failed' gpurun_out/r02_sanitizer_memcheck_tracker.log | tail -2
# sweep: feature noise x threshold (how selective the screen is) on cfg5
for fn in 0.01 0.02 0.03 0.04; do
  for thr in 0.7 1.0 1.2; do
    timeout 200 python bench.py --no-cpu-baseline --steps 6 --warmup 4 --feat-noise $fn --visual-threshold $thr \
      > gpurun_out/r02_sweep_fn${fn}_thr${thr}.json 2> /dev/null
  done
done
python - <<'PY'
import glob, json
rows = []
for fn in sorted(glob.glob("gpurun_out/r02_sweep_*.json")):
    try:
        d = json.loads(open(fn).read().strip().splitlines()[-1])
        rows.append({"feat_noise": d["config"].get("feat_noise"), "visual_threshold": d["config"]["option_overrides"]["visual_threshold"],
                     "ms_per_step": d["ms_per_step"], "value": d["value"], "stages_ms": d["stages_ms"],
                     "exact_fallback_scenes_per_step": d.get("exact_fallback_scenes_per_step")})
    except Exception as e:
        rows.append({"file": fn, "error": str(e)})
json.dump(rows, open("gpurun_out/r02_sweep_summary.json", "w"), indent=1)
for r in rows:
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k != "stages_ms"},
          {k: round(v, 3) for k, v in r.get("stages_ms", {}).items()})
PY
rm -f gpurun_out/r02_sweep_fn*.json
