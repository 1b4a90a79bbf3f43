mkdir -p gpurun_out
python bench.py --steps 8 --warmup 4 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_latest.json
python bench.py --impl reference --steps 2 --warmup 4 2>>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_ref_latest.json
tail -3 gpurun_out/bench_err.log
