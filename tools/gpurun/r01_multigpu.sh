#!/bin/bash
# multi-GPU bench exactly as the driver launches it (N = $1, default 2); also the reference arm under torchrun
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 10 --warmup 4 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"
tail -c 1500 gpurun_out/bench_n$N.json
grep -v "^$" gpurun_out/bench_n$N.err | tail -8
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N", "%.3e" % d["value"], round(d["ms_per_step"], 4), "e2e", "%.3e" % d["e2e"]["value"], round(d["e2e"]["ms_per_step"], 4))
except Exception as e:
    print("failed", e)
PY
