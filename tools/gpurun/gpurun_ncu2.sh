mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01_b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b2.log 2>&1
tail -2 gpurun_out/ncu_b2.log | cut -c1-300
