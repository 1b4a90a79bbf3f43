#!/usr/bin/env python
"""GPU0 -> GPU1 bandwidth three ways (what the ingest-rank scatter can expect): cudaMemcpyPeer inside one process, and -- under
torchrun with 2 ranks -- torch.distributed send/recv (NCCL) and this library's sb200_shard_scatter."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
N = 256 << 20


def single():
    a = torch.empty(N, dtype=torch.uint8, device="cuda:0")
    b = torch.empty(N, dtype=torch.uint8, device="cuda:1")
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t0 = time.perf_counter()
    for _ in range(10):
        b.copy_(a)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    dt = (time.perf_counter() - t0) / 10
    print("cudaMemcpyPeer 256 MiB: %.2f ms = %.1f GB/s" % (dt * 1e3, N / dt / 1e9))


def multi():
    import torch.distributed as dist

    import similari_b200.engine as eng

    rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    x = torch.empty(N, dtype=torch.uint8, device=dev)
    for _ in range(3):
        if rank == 0: dist.send(x, 1)
        else: dist.recv(x, 0)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        if rank == 0: dist.send(x, 1)
        else: dist.recv(x, 0)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    if rank == 0:
        print("torch.distributed send/recv 256 MiB: %.2f ms = %.1f GB/s" % (dt * 1e3, N / dt / 1e9))
    uid = [eng.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    c = eng.Comm(rank, 2, uid[0], local)
    rows = N // (24 + 2048) // 2
    rng = np.array([0, rows, 2 * rows], np.int32)
    allb = torch.zeros(2 * rows, 6, dtype=torch.float32, device=dev)
    allf = torch.zeros(2 * rows, 512, dtype=torch.float32, device=dev)
    myb = torch.zeros(rows, 6, dtype=torch.float32, device=dev)
    myf = torch.zeros(rows, 512, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream()
    for _ in range(3):
        c.scatter(0, rng, 512, allb.data_ptr() if rank == 0 else 0, allf.data_ptr() if rank == 0 else 0, myb.data_ptr(), myf.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        c.scatter(0, rng, 512, allb.data_ptr() if rank == 0 else 0, allf.data_ptr() if rank == 0 else 0, myb.data_ptr(), myf.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    if rank == 0:
        sent = rows * (24 + 2048)
        print("sb200_shard_scatter %d MiB to the peer: %.2f ms = %.1f GB/s" % (sent >> 20, dt * 1e3, sent / dt / 1e9))
    c.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    if "RANK" in os.environ:
        multi()
    else:
        single()
