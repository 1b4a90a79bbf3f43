import dataclasses, sys
import numpy as np
import similari_b200.engine as e
from similari_b200._lib import default_options
from similari_b200.workload import CONFIGS, Workload, tracker_options_for
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = dataclasses.replace(CONFIGS[name], n_scenes=S)
g = e.Tracker(tracker_options_for(name, default_options))
wl = Workload(cfg)
for fr in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    f = wl.next_frame()
    r = g.predict_batch(f["scene_ids"], f["det_offsets"], f["boxes"], features=f["features"])
    print(fr, len(r["ids"]), g.active_tracks(), flush=True)
