timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python bench.py --steps 8 --warmup 4 --no-cpu-baseline 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_latest.json; tail -3 gpurun_out/bench_err.log
