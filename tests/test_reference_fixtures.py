"""The reference's regression data (python/bugfixes) as committed fixtures, tests/golden/*.npz (inputs only: the
reference scripts assert invariants, not values).  CPU part: the oracle honours the invariants.  GPU part: the engine
gives the oracle's answer on the same real-world inputs."""
import pathlib

import numpy as np
import pytest

GOLDEN = pathlib.Path(__file__).parent / "golden"

# python/bugfixes/bug_vs_1/bug_visual_sort.py:18-33
BUG_VS_1 = dict(kind=2, positional_kind=0, max_idle_epochs=3, history_length=10, constraints=[(1, 1.0)],
                visual_kind=0, visual_threshold=1.0, feature_dim=512, visual_minimal_track_length=1,
                visual_minimal_area=5.0, visual_minimal_quality_use=0.45, visual_minimal_quality_collect=0.5,
                visual_max_observations=5, visual_min_votes=1)
# python/bugfixes/github-84.py:167-177
GITHUB_84 = dict(kind=0, positional_kind=1, iou_threshold=0.3, max_idle_epochs=5, history_length=10)


def _bug_vs_1_frames(seq):
    z = np.load(GOLDEN / "bug_vs_1.npz")
    return [(z[f"{seq}_{k}_boxes"], z[f"{seq}_{k}_features"], z[f"{seq}_{k}_quality"]) for k in range(2)]


def _run_visual(tracker, frames):
    out = []
    for boxes, feats, qual in frames:
        r = tracker.predict_batch([0], [0, len(boxes)], boxes, features=feats, quality=qual)
        out.append(r)
    return out


@pytest.mark.parametrize("seq", ["main", "fixed"])
def test_oracle_bug_vs_1_ids_unique_per_frame(oracle, seq):
    res = _run_visual(oracle.Tracker(oracle.make_options(**BUG_VS_1)), _bug_vs_1_frames(seq))
    for r in res:
        assert len(np.unique(r["ids"])) == len(r["ids"])          # bug_visual_sort.py:71-73
    assert np.isin(res[1]["ids"], res[0]["ids"]).sum() >= 2          # most people of frame 1 are found again in frame 2


def test_oracle_github_84_thin_oriented_boxes(oracle):
    z = np.load(GOLDEN / "github_84.npz")
    t = oracle.Tracker(oracle.make_options(**GITHUB_84))
    r1 = t.predict_batch([0], [0, len(z["boxes_1"])], z["boxes_1"])
    r2 = t.predict_batch([0], [0, len(z["boxes_2"])], z["boxes_2"])
    for r in (r1, r2):
        assert len(np.unique(r["ids"])) == len(r["ids"])
        assert np.all(np.isfinite(r["predicted"][:, [0, 1, 3, 4]])) and np.all(np.isfinite(r["observed"][:, [0, 1, 3, 4]]))
    costs = t.last_costs(0)
    assert costs.shape == (len(z["boxes_2"]), len(z["boxes_1"]))
    ok = costs[~np.isnan(costs)]
    assert np.all((ok >= 0.3) & (ok <= 1.0 + 1e-6))                 # IoU * confidence(1.0), thresholded at 0.3


@pytest.mark.gpu
@pytest.mark.parametrize("seq", ["main", "fixed"])
def test_gpu_bug_vs_1_matches_oracle(oracle, seq):
    import similari_b200.engine as eng
    from similari_b200._lib import default_options

    frames = _bug_vs_1_frames(seq)
    rg = _run_visual(eng.Tracker(default_options(**BUG_VS_1)), frames)
    ro = _run_visual(oracle.Tracker(oracle.make_options(**BUG_VS_1)), frames)
    for a, b in zip(rg, ro):
        for key in ("ids", "epochs", "lengths", "voting_types"):
            assert np.array_equal(a[key], b[key]), key
        assert np.array_equal(np.nan_to_num(a["predicted"], nan=-7.0), np.nan_to_num(b["predicted"], nan=-7.0))
        assert len(np.unique(a["ids"])) == len(a["ids"])


@pytest.mark.gpu
def test_gpu_github_84_matches_oracle(oracle):
    import similari_b200.engine as eng
    from similari_b200._lib import default_options

    z = np.load(GOLDEN / "github_84.npz")
    g, o = eng.Tracker(default_options(**GITHUB_84)), oracle.Tracker(oracle.make_options(**GITHUB_84))
    for name in ("boxes_1", "boxes_2"):
        a = g.predict_batch([0], [0, len(z[name])], z[name])
        b = o.predict_batch([0], [0, len(z[name])], z[name])
        for key in ("ids", "epochs", "lengths", "voting_types"):
            assert np.array_equal(a[key], b[key]), key
    cg, co = g.last_costs(0), o.last_costs(0)
    assert np.array_equal(np.isnan(cg), np.isnan(co))
    np.testing.assert_allclose(cg, co, rtol=0, atol=1e-6, equal_nan=True)   # oriented: device vs glibc sin/cos
