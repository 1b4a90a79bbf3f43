"""N > 1 path on CPU (gloo, world_size 2): the scene sharding logic of bench.py -- each rank owns a disjoint scene
range, generates its own frames, and the assigned track ids of all shards are gathered on every rank.  The GPU
engine is replaced by the oracle here (test infrastructure on both sides); what is exercised is the sharding,
the gather layout and the units / max-over-ranks reduction."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import oracle
    from similari_b200.workload import tracker_options_for

    n_sc = 3
    cfg, frames = bench.make_frames("cfg4", 3, scene_base=rank * n_sc, n_scenes_override=n_sc)
    t = oracle.Tracker(tracker_options_for("cfg4", oracle.make_options))
    max_total = 64
    units = 0
    ids_last = None
    for f in frames:
        assert set(map(int, f["scene_ids"])) == set(range(rank * n_sc, (rank + 1) * n_sc))   # disjoint scene shard
        offs = f["det_offsets"]
        # keep only the first 20 detections per scene to stay fast on CPU
        keep = np.concatenate([np.arange(offs[s], min(offs[s] + 20, offs[s + 1])) for s in range(n_sc)])
        new_offs = np.concatenate([[0], np.cumsum([min(20, offs[s + 1] - offs[s]) for s in range(n_sc)])]).astype(np.int32)
        n_before = np.array([len(t.scene_tracks(int(s))["ids"]) for s in f["scene_ids"]])
        r = t.predict_batch(f["scene_ids"], new_offs, f["boxes"][keep], want_boxes=False)
        units += int((np.diff(new_offs) * n_before).sum())
        ids_last = r["ids"]
    # gather the assigned ids of every shard (bench.py: all_gather_into_tensor on the device ids)
    pad = torch.zeros(max_total, dtype=torch.int64)
    pad[: len(ids_last)] = torch.from_numpy(ids_last.astype(np.int64))
    gathered = [torch.zeros(max_total, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, pad)
    un = torch.tensor([float(units)], dtype=torch.float64)
    dist.all_reduce(un, op=dist.ReduceOp.SUM)
    tm = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"),
            np.array([units, float(un[0]), float(tm[0])] + [int(g.sum()) for g in gathered], dtype=np.float64))
    dist.destroy_process_group()


def test_scene_sharding_two_ranks(tmp_path, oracle):
    world = 2
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert r0[1] == r1[1] == r0[0] + r1[0]          # units summed over ranks
    assert r0[2] == r1[2] == 2.0                     # time = max over ranks
    assert list(r0[3:]) == list(r1[3:])              # every rank sees the same gathered ids
    assert r0[3] > 0 and r0[4] > 0
