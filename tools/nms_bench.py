#!/usr/bin/env python
"""Times sb200_nms on the cfg5 NMS workload of SURVEY.md section 8d: 10 000 oriented boxes = 2 000 clusters x 5
near-duplicates (jitter 3 px / 0.03 rad) on 3840x2160, scores U(0,1), nms_threshold 0.8.  Host-pointer call
(H2D + kernels + D2H inside), best and median of `reps` runs; optional oracle check on a sub-sample."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from similari_b200.engine import nms_indices  # noqa: E402


def make_boxes(n_clusters=2000, dup=5, seed=0x5EED00A5):
    r = np.random.default_rng(seed)
    xc = r.uniform(0, 3840, n_clusters)
    yc = r.uniform(0, 2160, n_clusters)
    ang = r.uniform(-np.pi / 2, np.pi / 2, n_clusters)
    asp = r.uniform(0.3, 0.8, n_clusters)
    h = r.uniform(40, 160, n_clusters)
    b = np.empty((n_clusters, dup, 6), np.float32)
    b[..., 0] = xc[:, None] + r.normal(0, 3, (n_clusters, dup))
    b[..., 1] = yc[:, None] + r.normal(0, 3, (n_clusters, dup))
    b[..., 2] = ang[:, None] + r.normal(0, 0.03, (n_clusters, dup))
    b[..., 3] = asp[:, None]
    b[..., 4] = h[:, None]
    b[..., 5] = 1.0
    b = b.reshape(-1, 6)
    s = r.uniform(0, 1, len(b)).astype(np.float32)
    p = r.permutation(len(b))
    return np.ascontiguousarray(b[p]), s[p]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    boxes, scores = make_boxes()
    nms_indices(boxes, scores, 0.8)   # warm-up (context, allocations)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        keep = nms_indices(boxes, scores, 0.8)
        ts.append(time.perf_counter() - t0)
    out = {"workload": "NMS oriented 10k boxes (2000 clusters x 5), thr 0.8", "n": len(boxes), "kept": int(len(keep)),
           "ms_best": 1e3 * min(ts), "ms_median": 1e3 * float(np.median(ts)),
           "pair_tests_upper": len(boxes) * (len(boxes) - 1) // 2}
    if "--check" in sys.argv:
        import oracle

        t0 = time.perf_counter()
        ref = oracle.nms(boxes, scores, 0.8)
        out["oracle_ms"] = 1e3 * (time.perf_counter() - t0)
        out["identical_to_oracle"] = bool(np.array_equal(ref, keep))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
