#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'vis_screen|vis_refine' -s 8 -c 2 -f -o gpurun_out/prof_r01_f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_f.log | cut -c1-300
