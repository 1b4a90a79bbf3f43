#!/bin/bash
mkdir -p gpurun_out
SB200_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --visual-threshold max > gpurun_out/r02l_thrmax.json 2> gpurun_out/r02l_trace.err
grep "grows\|waits for\|store\|regrow" gpurun_out/r02l_trace.err | grep -v " 0 -> " | cut -c1-170 | tail -30
grep "predict:" gpurun_out/r02l_trace.err | awk '{ if ($4+0 > 5.0) print }' | tail -20
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02l_thrmax.json").read().strip().splitlines()[-1])
print("thrmax ms/step", round(d["ms_per_step"], 4), {k: round(v, 3) for k, v in d.get("host_ms_per_step", {}).items()})
PY
timeout 600 compute-sanitizer --tool initcheck --print-limit 30 python -m pytest tests/test_gpu_parity.py tests/test_gpu_own_area.py tests/test_gpu_api.py -m gpu -q -x > gpurun_out/r02_sanitizer_initcheck.log 2>&1
grep 'ERROR SUMMARY\|passed\|failed' gpurun_out/r02_sanitizer_initcheck.log | tail -2
grep "    at sb::" gpurun_out/r02_sanitizer_initcheck.log | sed 's/(.*)+0x[0-9a-f]*//' | sort | uniq -c | head -8
