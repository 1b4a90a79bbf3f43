#!/bin/bash
# round 2, final evidence pass: bench lines of every config, reference arm, launch list, steady-state full captures
# (summarised on the box: the .ncu-rep files stay there), host trace, sanitizer re-run
mkdir -p gpurun_out /tmp/ncu
show() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.4e" % d["value"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0) or 0, 3),
          {k: round(v, 3) for k, v in d.get("stages_ms", {}).items()}, "host", {k: round(v, 3) for k, v in d.get("host_ms_per_step", {}).items()},
          "roofline frac", round((d.get("roofline") or {}).get("frac", 0) or 0, 3), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_default.json 2> gpurun_out/r02_bench_n1_default.err
show gpurun_out/r02_bench_n1_default.json "default"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_reference_arm.json 2> /dev/null
show gpurun_out/r02_bench_n1_reference_arm.json "reference arm"
for cfg in 2 3 4; do
  timeout 300 python bench.py --no-cpu-baseline --config cfg$cfg --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg$cfg.json 2> /dev/null
  show gpurun_out/r02_bench_cfg$cfg.json cfg$cfg
done
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --visual-threshold max > gpurun_out/r02_bench_cfg5_thrmax.json 2>/dev/null
show gpurun_out/r02_bench_cfg5_thrmax.json "thr max (reference default)"
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --visual-threshold 10.0 > gpurun_out/r02_bench_cfg5_thr10.json 2>/dev/null
show gpurun_out/r02_bench_cfg5_thr10.json "thr 10.0 (reference bench)"
# host side: what the calling thread waits for
SB200_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 12 --warmup 5 > /dev/null 2> /tmp/ncu/trace.err
grep "predict waits\|predict:\|grows" /tmp/ncu/trace.err | grep -v " 0 -> " | tail -40 | cut -c1-170 > gpurun_out/r02_host_trace_cfg5.txt
tail -5 gpurun_out/r02_host_trace_cfg5.txt
# launch list of the bench command
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_cfg5_final.csv python bench.py --no-cpu-baseline --steps 2 --warmup 6 > /tmp/ncu/ncu_bench.log 2>&1
cap() {  # name, args...
  local name=$1; shift
  timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o /tmp/ncu/$name \
    python tools/profile_frame.py "$@" > /tmp/ncu/$name.log 2>&1
  tail -1 /tmp/ncu/$name.log | cut -c1-300
  python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep > gpurun_out/r02_ncu_${name}_summary.txt 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > /tmp/ncu/$name.raw.csv 2>/dev/null
  python tools/ncu_pick.py /tmp/ncu/$name.raw.csv > gpurun_out/r02_ncu_${name}_metrics.csv 2>/dev/null
}
cap frame_cfg5 cfg5 9
cap frame_cfg5_dense cfg5 9 --visual-threshold max
cap frame_cfg4 cfg4 9
cap frame_cfg2 cfg2 9
ncu -i /tmp/ncu/frame_cfg5_dense.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:vis_wsum > /tmp/ncu/wsum_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/ncu/wsum_src.csv 30 > gpurun_out/r02_ncu_wsum_lines.txt 2>&1
ncu -i /tmp/ncu/frame_cfg5_dense.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:vis_dense_select > /tmp/ncu/sel_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/ncu/sel_src.csv 20 > gpurun_out/r02_ncu_select_lines.txt 2>&1
ncu -i /tmp/ncu/frame_cfg4.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:pos_scan > /tmp/ncu/pos_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/ncu/pos_src.csv 25 > gpurun_out/r02_ncu_pos_scan_maha_lines.txt 2>&1
ncu -i /tmp/ncu/frame_cfg2.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:pos_scan > /tmp/ncu/pos2_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/ncu/pos2_src.csv 25 > gpurun_out/r02_ncu_pos_scan_iou_lines.txt 2>&1
ncu -i /tmp/ncu/frame_cfg5.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:apply_kernel > /tmp/ncu/apply_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/ncu/apply_src.csv 20 > gpurun_out/r02_ncu_apply_lines.txt 2>&1
# sanitizer
SEL='tests/test_gpu_parity.py tests/test_gpu_own_area.py tests/test_gpu_api.py'
for tool in memcheck racecheck initcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 30 python -m pytest $SEL -m gpu -q -x > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep 'ERROR SUMMARY\|passed\|failed' gpurun_out/r02_sanitizer_$tool.log | tail -2 | tr '\n' ' ')"
  grep "    at sb::" gpurun_out/r02_sanitizer_$tool.log | sed 's/(.*)+0x[0-9a-f]*//' | sort | uniq -c | head -8
done
timeout 1200 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_tracker.py -m gpu -q -x \
  -k "frames_in_flight or ragged or history or gated_pair or (dense_tensor_core and 3-2-0-3.4 and 3])" > gpurun_out/r02_sanitizer_memcheck_tracker.log 2>&1
echo "== memcheck tracker: $(grep 'ERROR SUMMARY\|passed\|failed' gpurun_out/r02_sanitizer_memcheck_tracker.log | tail -2 | tr '\n' ' ')"
du -sh gpurun_out
