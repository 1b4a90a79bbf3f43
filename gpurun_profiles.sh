#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_cfg5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'vis_screen|vis_refine|feat_store|pos_scan|voting_sparse' -s 20 -c 5 -f -o gpurun_out/prof_r01_h python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_h.log 2>&1
tail -2 gpurun_out/ncu_h.log | cut -c1-200
