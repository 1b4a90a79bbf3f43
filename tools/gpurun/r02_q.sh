#!/bin/bash
# round 2: refresh of the bench lines / launch list / host trace with the final code (full captures: r02_o.sh)
mkdir -p gpurun_out /tmp/ncu
show() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.4e" % d["value"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0) or 0, 3),
          {k: round(v, 3) for k, v in d.get("stages_ms", {}).items()}, "host", {k: round(v, 3) for k, v in d.get("host_ms_per_step", {}).items()},
          "roofline frac", round((d.get("roofline") or {}).get("frac", 0) or 0, 3))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02q_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_default.json 2> gpurun_out/r02_bench_n1_default.err
show gpurun_out/r02_bench_n1_default.json "default"
for cfg in 2 3 4; do
  timeout 300 python bench.py --no-cpu-baseline --config cfg$cfg --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg$cfg.json 2> /dev/null
  show gpurun_out/r02_bench_cfg$cfg.json cfg$cfg
done
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --visual-threshold max > gpurun_out/r02_bench_cfg5_thrmax.json 2>/dev/null
show gpurun_out/r02_bench_cfg5_thrmax.json "thr max"
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --visual-threshold 10.0 > gpurun_out/r02_bench_cfg5_thr10.json 2>/dev/null
show gpurun_out/r02_bench_cfg5_thr10.json "thr 10.0"
SB200_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 12 --warmup 5 > /dev/null 2> /tmp/ncu/trace.err
grep "predict waits\|predict:\|grows" /tmp/ncu/trace.err | grep -v " 0 -> " | tail -40 | cut -c1-170 > gpurun_out/r02_host_trace_cfg5.txt
tail -3 gpurun_out/r02_host_trace_cfg5.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_cfg5_final.csv python bench.py --no-cpu-baseline --steps 2 --warmup 6 > /tmp/ncu/ncu_bench.log 2>&1
tail -2 gpurun_out/r02_launches_cfg5_final.csv | cut -c1-200
