"""The reference-named Python classes (similari_b200.api) on the GPU: the reference's own usage sequences."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sort_api_sequence():
    # src/trackers/sort/simple_api.rs:280-342 through the PyO3-named classes
    from similari_b200.api import BoundingBox, PositionalMetricType, Sort

    t = Sort(shards=1, bbox_history=10, max_idle_epochs=2, method=PositionalMetricType.iou(0.3), min_confidence=0.05)
    v = t.predict([(BoundingBox(0.0, 0.0, 10.0, 20.0).as_xyaah(), None)])
    assert len(v) == 1 and v[0].length == 1 and v[0].epoch == 1 and v[0].custom_object_id is None
    tid = v[0].id
    v = t.predict([(BoundingBox(0.1, 0.1, 10.1, 20.0).as_xyaah(), 2)])
    assert v[0].id == tid and v[0].length == 2 and v[0].custom_object_id == 2
    v = t.predict([(BoundingBox(10.1, 10.1, 10.1, 20.0).as_xyaah(), 3)])
    assert v[0].id != tid
    assert t.wasted() == [] and t.current_epoch() == 3
    t.predict([])
    t.predict([])
    w = t.wasted()
    assert [x.id for x in w] == [tid]


def test_batch_sort_api():
    from similari_b200.api import BatchSort, BoundingBox, PositionalMetricType, SortPredictionBatchRequest

    t = BatchSort(method=PositionalMetricType.iou(0.3))
    req = SortPredictionBatchRequest()
    req.add(0, BoundingBox(0.0, 0.0, 5.0, 7.0).as_xyaah(), 11)
    req.add(1, BoundingBox(0.0, 0.0, 5.0, 7.0).as_xyaah(), 12)
    res = t.predict(req)
    assert res.batch_size() == 2
    got = dict(res.get() for _ in range(2))
    assert got[0][0].custom_object_id == 11 and got[1][0].custom_object_id == 12 and got[0][0].id != got[1][0].id


def test_visual_sort_api_and_nms():
    from similari_b200.api import (BoundingBox, PositionalMetricType, VisualSort, VisualSortMetricType,
                                   VisualSortObservation, VisualSortObservationSet, VisualSortOptions, VotingType, nms)

    o = VisualSortOptions()
    o.max_idle_epochs(3)
    o.visual_metric(VisualSortMetricType.euclidean(1.0))
    o.positional_metric(PositionalMetricType.maha())
    o.visual_minimal_track_length(2)
    o.visual_max_observations(3)
    o.visual_min_votes(2)
    t = VisualSort(1, o)

    def step(feat, box):
        s = VisualSortObservationSet()
        s.add(VisualSortObservation(feat, 0.9, BoundingBox(*box).as_xyaah(), None))
        return t.predict_with_scene(10, s)[0]

    a = step([1.0, 1.0], (1.0, 1.0, 3.0, 5.0))
    b = step([0.95, 0.95], (1.1, 1.1, 3.05, 5.01))
    c = step([0.97, 0.97], (1.15, 1.2, 3.1, 5.05))
    assert a.id == b.id == c.id and a.voting_type == VotingType.Positional and c.voting_type == VotingType.Visual

    bbox1 = (BoundingBox(10.0, 11.0, 3.0, 3.8).as_xyaah(), 1.0)
    bbox2 = (BoundingBox(10.3, 11.1, 2.9, 3.9).as_xyaah(), 0.9)
    res = nms([bbox2, bbox1], nms_threshold=0.7, score_threshold=0.0)  # src/utils/nms/nms_py.rs:23-39
    assert len(res) == 1 and abs(float(res[0].xc) - 11.5) < 1e-5


def test_visual_sort_first_frames_without_features():
    """VisualSortObservation.feature is an Option (src/trackers/visual_sort.rs:42-55): a tracker may be fed feature-less
    frames before the first ReID vector arrives.  The track created by those frames must survive the switch to the real
    feature length (round 1 pinned a provisional length of 8 and failed on the first 128-d feature)."""
    from similari_b200.api import (BoundingBox, PositionalMetricType, VisualSort, VisualSortMetricType,
                                   VisualSortObservation, VisualSortObservationSet, VisualSortOptions, VotingType)

    o = VisualSortOptions()
    o.max_idle_epochs(3)
    o.visual_metric(VisualSortMetricType.euclidean(1.0))
    o.positional_metric(PositionalMetricType.maha())
    o.visual_minimal_track_length(2)
    o.visual_max_observations(3)
    o.visual_min_votes(2)
    t = VisualSort(1, o)
    rng = np.random.default_rng(4)
    v = rng.standard_normal(128).astype(np.float32)
    v /= np.linalg.norm(v)

    def step(feat, box):
        s = VisualSortObservationSet()
        s.add(VisualSortObservation(None if feat is None else list(map(float, feat)), 0.9, BoundingBox(*box).as_xyaah(), None))
        return t.predict_with_scene(3, s)[0]

    a = step(None, (1.0, 1.0, 3.0, 5.0))
    b = step(None, (1.05, 1.05, 3.0, 5.0))
    c = step(v, (1.1, 1.1, 3.02, 5.0))                       # first feature: 128-d
    d = step(v + 0.01, (1.15, 1.15, 3.02, 5.0))
    e = step(v - 0.01, (1.2, 1.2, 3.0, 5.0))
    assert a.id == b.id == c.id == d.id == e.id and e.length == 5
    assert c.voting_type == VotingType.Positional and e.voting_type == VotingType.Visual
    with pytest.raises(Exception):
        step(np.ones(64, np.float32), (1.2, 1.2, 3.0, 5.0))  # a second feature length is an error, as in the reference
