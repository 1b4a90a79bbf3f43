#!/usr/bin/env python
"""One steady-state frame of a BASELINE config between cudaProfilerStart / cudaProfilerStop, inputs resident in HBM:
    ncu --profile-from-start off --set full --import-source on -o out python tools/profile_frame.py cfg5 [warm] [--visual-threshold max]
captures every kernel of exactly that frame (the frame bench.py times), nothing of the warm-up."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    import similari_b200.engine as eng
    from similari_b200._lib import default_options
    from similari_b200.workload import CONFIGS, Workload, tracker_options_for

    name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    warm = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 9
    over = {}
    if "--visual-threshold" in sys.argv:
        v = sys.argv[sys.argv.index("--visual-threshold") + 1]
        over["visual_threshold"] = 3.402823466e38 if v == "max" else float(v)
    cfg = CONFIGS[name]
    dev = torch.device("cuda", 0)
    t = eng.Tracker(tracker_options_for(name, default_options, max_scenes_hint=cfg.n_scenes,
                                        max_tracks_per_scene_hint=4 * cfg.n_objects, max_dets_per_scene_hint=cfg.n_objects, **over))
    t.set_stream(torch.cuda.current_stream().cuda_stream)
    wl = Workload(cfg)
    rt = torch.cuda.cudart()
    n = cfg.n_scenes * cfg.n_objects
    d_ids = torch.zeros(n, dtype=torch.int64, device=dev)
    for fr in range(warm + 1):
        f = wl.next_frame()
        db = torch.from_numpy(np.ascontiguousarray(f["boxes"])).to(dev)
        df = torch.from_numpy(f["features"]).to(dev) if f["features"] is not None else None
        torch.cuda.synchronize()
        if fr == warm:
            t.sync()
            rt.cudaProfilerStart()
        t.predict_batch_device(f["scene_ids"], f["det_offsets"], db.data_ptr(), df.data_ptr() if df is not None else 0,
                               d_ids=d_ids.data_ptr())
        t.sync()
        if fr == warm:
            rt.cudaProfilerStop()
    wc = t.work_counters()
    print({k: wc[k] for k in ("frames", "tc_frames", "dense_fallback_scenes")}, "last frame stage ms:", t.last_stage_ms(), t.last_kernel_ms())


if __name__ == "__main__":
    main()
