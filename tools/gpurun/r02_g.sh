#!/bin/bash
# round 2: dense-path trace + on-box summaries of the steady-state full captures (the .ncu-rep files stay on the box)
mkdir -p gpurun_out /tmp/ncu
SB200_TRACE=1 timeout 300 python bench.py --visual-threshold max --no-cpu-baseline --steps 10 > gpurun_out/r02g_bench_thrmax.json 2> /tmp/ncu/thrmax.err
grep "predict:" /tmp/ncu/thrmax.err | tail -24 | cut -c1-160
grep -c "absorbed" /tmp/ncu/thrmax.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02g_bench_thrmax.json").read().strip().splitlines()[-1])
print("thrmax", "%.4e" % d["value"], "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 3), d["clocks"], {k: round(v, 4) for k, v in d["stages_ms"].items()})
PY
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
# source-level view of the two hot kernels of the dense path and of the positional scan
ncu -i /tmp/ncu/frame_cfg5_dense.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:vis_wsum > /tmp/ncu/wsum_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/ncu/wsum_src.csv 30 > gpurun_out/r02_ncu_wsum_lines.txt 2>&1
ncu -i /tmp/ncu/frame_cfg5_dense.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:vis_dense_select > /tmp/ncu/sel_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/ncu/sel_src.csv 20 > gpurun_out/r02_ncu_select_lines.txt 2>&1
ncu -i /tmp/ncu/frame_cfg4.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:pos_scan > /tmp/ncu/pos_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/ncu/pos_src.csv 25 > gpurun_out/r02_ncu_pos_scan_maha_lines.txt 2>&1
ncu -i /tmp/ncu/frame_cfg2.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:pos_scan > /tmp/ncu/pos2_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/ncu/pos2_src.csv 25 > gpurun_out/r02_ncu_pos_scan_iou_lines.txt 2>&1
ls -la gpurun_out/ | head -30
du -sh gpurun_out
