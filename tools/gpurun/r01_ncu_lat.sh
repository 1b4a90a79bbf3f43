#!/bin/bash
# full capture of the latency-bound per-scene kernels (positional scan, apply, end-of-frame sweep, sparse voting)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pos_scan|apply_kernel|waste_kernel|voting_sparse' -s 24 -c 4 -f -o gpurun_out/prof_r01_lat python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lat.log 2>&1
tail -3 gpurun_out/ncu_lat.log | cut -c1-300
ls -la gpurun_out/prof_r01_lat.ncu-rep
