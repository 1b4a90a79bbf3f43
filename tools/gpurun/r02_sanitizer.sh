#!/bin/bash
# compute-sanitizer passes over the small GPU parity tests (memcheck found the one out-of-bounds read of round 1 in
# seconds; racecheck / initcheck / synccheck have not been run yet).  ~10-50x slower than a plain run: small tests only.
mkdir -p gpurun_out
SEL='tests/test_gpu_parity.py tests/test_gpu_own_area.py tests/test_gpu_api.py'
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest $SEL -m gpu -q -x \
    > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -c 'Invalid\|Race reported\|Uninitialized\|Barrier error' gpurun_out/sanitizer_$tool.log) finding line(s)"
  grep -m3 -A3 'Invalid\|Race reported\|Uninitialized\|Barrier error' gpurun_out/sanitizer_$tool.log | cut -c1-200
  grep 'ERROR SUMMARY\|passed\|failed' gpurun_out/sanitizer_$tool.log | tail -2
done
