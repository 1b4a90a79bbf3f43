#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=INFO
export NCCL_DEBUG_SUBSYS=INIT,P2P,SHM,NET,GRAPH
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 \
  bench.py --gpus 2 --steps 4 --warmup 3 --scenes 32 > gpurun_out/r02_n2dbg.json 2> gpurun_out/r02_n2dbg.err
grep -i "via\|P2P\|SHM\|NVLS\|channels\|cuMem\|IPC\|transport" gpurun_out/r02_n2dbg.err | sort | uniq -c | sort -rn | head -40 | cut -c1-220
nvidia-smi topo -m | head -6
ls -la /dev/shm | head -5; df -h /dev/shm | tail -1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_n2dbg.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("scatter_ingest"))[:600])
PY
