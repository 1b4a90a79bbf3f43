#!/bin/bash
# cluster-multicast screen kernel: parity first (bounded), then bench both variants
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tracker.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -15
timeout 300 python bench.py --steps 8 --warmup 4 > gpurun_out/bench_cluster.json 2> gpurun_out/bench_cluster.err; tail -c 3000 gpurun_out/bench_cluster.json
SB200_SCREEN_SINGLE=1 timeout 300 python bench.py --steps 8 --warmup 4 > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; tail -c 1200 gpurun_out/bench_single.json
