mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01_a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --scenes 64 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vis_cost -s 3 -c 1 -o gpurun_out/prof_vis_simt python bench.py --steps 1 --warmup 3 --no-cpu-baseline --scenes 64 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out
