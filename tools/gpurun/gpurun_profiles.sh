#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'vis_screen|vis_refine' -s 12 -c 2 -f -o gpurun_out/prof_r01_i python bench.py --steps 2 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_i.log 2>&1
tail -2 gpurun_out/ncu_i.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_cfg5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
