#!/bin/bash
# four GPUs, launched exactly as the driver does: local-ingest headline + the ingest-rank scatter arm (sb200_shard_*)
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
tail -5 gpurun_out/r02_bench_n4.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n4.json").read().strip().splitlines()[-1])
    print("N=4 value %.4e" % d["value"], "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "launches/step", d["gpu_launches_per_step"])
    print("scatter arm:", json.dumps(d.get("scatter_ingest"))[:900])
except Exception as e:
    print("failed", e)
PY
