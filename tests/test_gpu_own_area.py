"""GPU parity of the own-area shares (kernels_own.cu) with the oracle: the stateless operator and the visual trackers
that derive the shares themselves when an own-area threshold is set (src/trackers/visual_sort/simple_api.rs:110-127)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import similari_b200.engine as e
    from similari_b200._lib import lib

    if lib().sb200_device_count() <= 0:
        pytest.fail("no CUDA device: the product has no CPU path")
    return e


def ltwh(l, t, w, h):
    return [l + w / 2, t + h / 2, np.nan, w / h, h, 1.0]


def random_boxes(r, n, oriented, span):
    b = np.zeros((n, 6), np.float32)
    b[:, 0] = r.uniform(0, span, n)
    b[:, 1] = r.uniform(0, span, n)
    b[:, 2] = r.uniform(-1.5, 1.5, n) if oriented else np.nan
    b[:, 3] = r.uniform(0.3, 0.8, n)
    b[:, 4] = r.uniform(40, 160, n)
    b[:, 5] = 1
    return b


@pytest.mark.parametrize("oriented", [False, True])
def test_operator_matches_oracle(eng, oracle, oriented):
    r = np.random.default_rng(31 + oriented)
    for n, span in ((1, 100.0), (2, 60.0), (7, 120.0), (33, 300.0), (64, 400.0), (300, 1500.0), (512, 3000.0)):
        b = random_boxes(r, n, oriented, span)
        got, ref = eng.own_area_shares(b), oracle.own_area_shares(b)
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-6)
    assert 0.05 < float((ref < 0.999).mean())


def test_operator_coincident_outlines_and_reference_case(eng, oracle):
    cases = [
        [ltwh(0, 0, 10, 10), ltwh(5, 5, 10, 10), ltwh(10, 10, 10, 10)],      # bbox_own_areas.rs:50-79 -> .75 .5 .75
        [ltwh(0, 0, 10, 10), ltwh(0, 0, 10, 10)],
        [ltwh(0, 0, 10, 10)] * 3,
        [ltwh(0, 0, 10, 10), ltwh(2, 2, 4, 4)],
        [ltwh(0, 0, 10, 10), ltwh(10, 0, 10, 10)],
        [ltwh(0, 0, 10, 10), ltwh(5, 0, 10, 10), ltwh(5, 0, 10, 10)],
        [ltwh(0, 0, 10, 10), ltwh(0, 0, 5, 10), ltwh(5, 0, 5, 10)],
        [ltwh(0, 4, 10, 2), ltwh(4, 0, 2, 10)],
    ]
    for boxes in cases:
        np.testing.assert_allclose(eng.own_area_shares(boxes), oracle.own_area_shares(boxes), rtol=0, atol=1e-6)
    got = eng.own_area_shares(cases[0])
    assert np.all(np.abs(got - np.array([0.75, 0.5, 0.75], np.float32)) < 1e-5)


def test_operator_crowds_take_the_second_pass(eng, oracle):
    """More than 32 boxes overlapping one box (the warp kernel's on-chip list) go through the CTA-per-box second pass; the
    reference has no limit (src/utils/clipping/bbox_own_areas.rs:8-46)."""
    got = eng.own_area_shares([ltwh(0, 0, 10, 10)] * 40)         # 39 coincident boxes overlap each box
    np.testing.assert_allclose(got, oracle.own_area_shares([ltwh(0, 0, 10, 10)] * 40), rtol=0, atol=1e-6)
    r = np.random.default_rng(5)
    for oriented in (False, True):
        for n, span in ((60, 120.0), (150, 260.0), (400, 500.0)):   # dense crowds: 40..150 boxes overlap each box
            b = random_boxes(r, n, oriented, span)
            got, ref = eng.own_area_shares(b), oracle.own_area_shares(b)
            np.testing.assert_allclose(got, ref, rtol=0, atol=1e-6)


def test_visual_tracker_in_a_crowd_never_errors(eng, oracle):
    """The tracker path of the same: a crowded scene (every detection overlapped by > 32 others) inside a batch; predict
    succeeds and matches the oracle (round 1 returned SB200_ERR_CAPACITY after the store had been advanced)."""
    from similari_b200._lib import default_options

    r = np.random.default_rng(8)
    kw = dict(kind=3, positional_kind=1, iou_threshold=0.3, max_idle_epochs=3, visual_kind=0, visual_threshold=0.7,
              feature_dim=32, visual_max_observations=3, visual_min_votes=1, visual_minimal_track_length=1,
              visual_minimal_own_area_percentage_use=0.05, visual_minimal_own_area_percentage_collect=0.1)
    g, o = eng.Tracker(default_options(**kw)), oracle.Tracker(oracle.make_options(**kw))
    n_obj = 80
    base = [random_boxes(r, n_obj, False, 150.0), random_boxes(r, n_obj, True, 1500.0)]   # scene 0 is the crowd
    cent = r.standard_normal((2, n_obj, 32)).astype(np.float32)
    offs = np.arange(3, dtype=np.int32) * n_obj
    for fr in range(4):
        boxes = np.concatenate(base).copy()
        boxes[:, :2] += r.normal(0, 1.0, (len(boxes), 2)).astype(np.float32)
        feats = (cent + 0.01 * r.standard_normal(cent.shape).astype(np.float32)).reshape(-1, 32)
        rg = g.predict_batch(np.arange(2), offs, boxes, features=feats)
        ro = o.predict_batch(np.arange(2), offs, boxes, features=feats)
        for key in ("ids", "epochs", "lengths", "voting_types"):
            assert np.array_equal(rg[key], ro[key]), (fr, key)
        for s in range(2):
            assert np.array_equal(g.scene_tracks(s)["feat_counts"], o.scene_tracks(s)["feat_counts"]), (fr, s)


def test_visual_tracker_derives_shares_like_the_oracle(eng, oracle):
    from similari_b200._lib import default_options

    r = np.random.default_rng(3)
    kw = dict(kind=3, positional_kind=1, iou_threshold=0.3, max_idle_epochs=3, visual_kind=0, visual_threshold=0.7,
              feature_dim=32, visual_max_observations=3, visual_min_votes=1, visual_minimal_track_length=1,
              visual_minimal_own_area_percentage_use=0.6, visual_minimal_own_area_percentage_collect=0.7)
    g, o = eng.Tracker(default_options(**kw)), oracle.Tracker(oracle.make_options(**kw))
    n_sc, n_obj = 3, 24
    base = [random_boxes(r, n_obj, s % 2 == 1, 260.0) for s in range(n_sc)]
    cent = r.standard_normal((n_sc, n_obj, 32)).astype(np.float32)
    offs = np.arange(n_sc + 1, dtype=np.int32) * n_obj
    for fr in range(6):
        boxes = np.concatenate(base).copy()
        boxes[:, :2] += r.normal(0, 1.0, (len(boxes), 2)).astype(np.float32)
        feats = (cent + 0.01 * r.standard_normal(cent.shape).astype(np.float32)).reshape(-1, 32)
        rg = g.predict_batch(np.arange(n_sc), offs, boxes, features=feats)
        ro = o.predict_batch(np.arange(n_sc), offs, boxes, features=feats)
        for key in ("ids", "epochs", "lengths", "voting_types"):
            assert np.array_equal(rg[key], ro[key]), (fr, key)
        for s in range(n_sc):
            assert np.array_equal(g.scene_tracks(s)["feat_counts"], o.scene_tracks(s)["feat_counts"]), (fr, s)
    own = oracle.own_area_shares(np.concatenate(base)[:n_obj])
    assert own.min() < 0.6 < own.max()          # the thresholds did separate covered from free detections
