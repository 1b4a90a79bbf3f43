#!/usr/bin/env python
"""Runs a few frames of cfg5 with the reference's default visual metric (Euclidean(f32::MAX)) on the dense tensor-core
path with SB200_TRACE=1, so that the per-frame diagnostics (fallback reasons, list lengths) land on stderr."""
import dataclasses
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SB200_TRACE"] = "1"


def main():
    import similari_b200.engine as eng
    from similari_b200._lib import default_options
    from similari_b200.workload import CONFIGS, Workload, tracker_options_for

    n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    cfg = dataclasses.replace(CONFIGS["cfg5"], n_scenes=n_scenes)
    g = eng.Tracker(tracker_options_for("cfg5", default_options, visual_threshold=float(np.finfo(np.float32).max)))
    wl = Workload(cfg)
    for fr in range(6):
        f = wl.next_frame()
        print(f"--- frame {fr}", file=sys.stderr)
        g.predict_batch(f["scene_ids"], f["det_offsets"], f["boxes"], features=f["features"])
    print(g.work_counters(), file=sys.stderr)


if __name__ == "__main__":
    main()
