"""BASELINE.json configs at (or near) their full per-scene sizes: GPU engine vs the oracle, plus size-independent
properties where the oracle would take too long."""
import dataclasses

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import similari_b200.engine as e
    from similari_b200._lib import lib

    if lib().sb200_device_count() <= 0:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    return e


def _run(eng, oracle, name, n_scenes, frames, threads=16, **cfg_over):
    from similari_b200._lib import default_options
    from similari_b200.workload import CONFIGS, Workload, tracker_options_for

    cfg = dataclasses.replace(CONFIGS[name], n_scenes=n_scenes, **cfg_over)
    g = eng.Tracker(tracker_options_for(name, default_options))
    o = oracle.Tracker(tracker_options_for(name, oracle.make_options), threads=threads)
    wl = Workload(cfg)
    for fr in range(frames):
        f = wl.next_frame()
        rg = g.predict_batch(f["scene_ids"], f["det_offsets"], f["boxes"], features=f["features"])
        ro = o.predict_batch(f["scene_ids"], f["det_offsets"], f["boxes"], features=f["features"])
        for key in ("ids", "epochs", "lengths", "voting_types"):
            assert np.array_equal(rg[key], ro[key]), (name, fr, key, int((rg[key] != ro[key]).sum()))
    return g, o


def test_cfg1_sort_iou_100x100_sparse_bench_workload(eng, oracle):
    """benches/simple_sort_iou_tracker.rs: 100 objects 1000 px apart, constraints (1, 1.0), Sort IoU(0.3)."""
    from similari_b200._lib import default_options
    from similari_b200.workload import tracker_options_for

    g = eng.Tracker(tracker_options_for("cfg1", default_options))
    o = oracle.Tracker(tracker_options_for("cfg1", oracle.make_options))
    rng = np.random.default_rng(1)
    x = 1000.0 * np.arange(100, dtype=np.float32)
    y = x.copy()
    w = np.full(100, 50.0, np.float32)
    h = np.full(100, 50.0, np.float32)
    for it in range(12):
        x += rng.uniform(-1, 1, 100).astype(np.float32)
        y += rng.uniform(-1, 1, 100).astype(np.float32)
        w = np.maximum(w + rng.uniform(-0.001, 0.001, 100).astype(np.float32), 1.0)
        h = np.maximum(h + rng.uniform(-0.001, 0.001, 100).astype(np.float32), 1.0)
        boxes = np.stack([x + w / 2, y + h / 2, np.full(100, np.nan, np.float32), w / h, h, np.ones(100, np.float32)], 1)
        rg = g.predict_batch([0], [0, 100], boxes)
        ro = o.predict_batch([0], [0, 100], boxes)
        assert np.array_equal(rg["ids"], ro["ids"]) and np.array_equal(rg["lengths"], ro["lengths"])
    assert g.active_tracks() == 100 and np.all(rg["lengths"] == 12)      # the bench's own assertions
    g.skip_epochs(2)
    assert len(g.wasted()["ids"]) == 100


def test_cfg2_batchsort_iou_64x256x256_matches_oracle(eng, oracle):
    _run(eng, oracle, "cfg2", 64, 5)


def test_cfg3_visualsort_cosine_1024x1024x512_matches_oracle(eng, oracle):
    _run(eng, oracle, "cfg3", 1, 4)


def test_cfg4_batchsort_maha_oriented_512x512_matches_oracle(eng, oracle):
    _run(eng, oracle, "cfg4", 16, 5)


def test_cfg5_batchvisualsort_512x512x512_matches_oracle(eng, oracle):
    _run(eng, oracle, "cfg5", 8, 5)


def test_cfg5_full_size_properties(eng):
    """256 scenes x 512 x 512 x 512-d: properties that do not need the oracle."""
    from similari_b200._lib import default_options
    from similari_b200.workload import CONFIGS, Workload, tracker_options_for

    cfg = CONFIGS["cfg5"]
    t = eng.Tracker(tracker_options_for("cfg5", default_options, max_scenes_hint=cfg.n_scenes,
                                        max_tracks_per_scene_hint=1200, max_dets_per_scene_hint=cfg.n_objects))
    wl = Workload(cfg)
    prev = None
    seen = np.empty(0, np.uint64)
    for fr in range(4):
        f = wl.next_frame()
        r = t.predict_batch(f["scene_ids"], f["det_offsets"], f["boxes"], features=f["features"],
                            want=("ids", "epochs", "lengths", "voting_types"))
        offs = f["det_offsets"]
        assert np.all(r["epochs"] == fr + 1)
        # a track id is handed out at most once per frame (python/bugfixes/bug_vs_1: "ids unique per frame")
        assert len(np.unique(r["ids"])) == len(r["ids"])
        if prev is not None:
            cont = np.isin(r["ids"], prev)
            assert cont.mean() > 0.85                       # ~90 % of the detections continue a track
            assert np.all(r["lengths"][~np.isin(r["ids"], seen)] == 1)   # never-seen ids start at length 1
            if fr >= 2:
                assert (r["voting_types"][cont] == 0).mean() > 0.9   # with >= 2 stored features the match is Visual
        prev = r["ids"].copy()
        seen = np.union1d(seen, prev)
    assert t.active_tracks() == sum(int(x) for x in t.scene_track_counts(f["scene_ids"]))


def test_nms_10k_oriented_boxes_matches_oracle(eng, oracle):
    """cfg5's NMS part: 2000 clusters x 5 near-duplicates on 3840 x 2160, threshold 0.8."""
    rng = np.random.default_rng(5)
    base = np.empty((2000, 6), np.float32)
    base[:, 0] = rng.uniform(0, 3840, 2000)
    base[:, 1] = rng.uniform(0, 2160, 2000)
    base[:, 2] = rng.uniform(-1.5, 1.5, 2000)
    base[:, 3] = rng.uniform(0.3, 0.8, 2000)
    base[:, 4] = rng.uniform(40, 160, 2000)
    base[:, 5] = 1.0
    boxes = np.repeat(base, 5, axis=0)
    boxes[:, :2] += rng.normal(0, 3, (10000, 2)).astype(np.float32)
    boxes[:, 2] += rng.normal(0, 0.03, 10000).astype(np.float32)
    scores = rng.uniform(0, 1, 10000).astype(np.float32)
    ref = oracle.nms(boxes, scores, 0.8)
    got = eng.nms_indices(boxes, scores, 0.8)
    assert list(ref) == list(got) and 2000 <= len(got) < 10000
