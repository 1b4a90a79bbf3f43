#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01_g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_g.log 2>&1
tail -1 gpurun_out/ncu_g.log | cut -c1-200
