#!/bin/bash
# Round-1 final ncu evidence: launch list of one bench command + full capture of the two dominant visual-cost kernels.
# A number printed by a run under ncu is never a bench value.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r01_launches_cfg5_v2.csv \
  python bench.py --steps 3 --warmup 6 --no-cpu-baseline > gpurun_out/ncu_list_v2.log 2>&1
tail -1 gpurun_out/ncu_list_v2.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'vis_screen|vis_refine' -s 16 -c 2 -f -o gpurun_out/prof_r01_v2 \
  python bench.py --steps 3 --warmup 6 --no-cpu-baseline > gpurun_out/ncu_full_v2.log 2>&1
tail -2 gpurun_out/ncu_full_v2.log | cut -c1-200
ls -la gpurun_out/prof_r01_v2.ncu-rep gpurun_out/r01_launches_cfg5_v2.csv
