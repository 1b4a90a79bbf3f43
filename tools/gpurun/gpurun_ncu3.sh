mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"vis_screen|pos_cost" -s 6 -c 2 -o gpurun_out/prof_r01_c python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
ls -la gpurun_out
