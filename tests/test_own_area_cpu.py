"""exclusively_owned_areas_normalized_shares (src/utils/clipping/bbox_own_areas.rs:8-46): the oracle against the
reference's own test, against an independent inclusion-exclusion evaluation with a plain Sutherland-Hodgman clipper,
on degenerate (coincident-outline) configurations, and against the product's arithmetic (sb_own_area.cuh, host build)."""
import ctypes as C
import itertools
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def ltwh(l, t, w, h):
    return [l + w / 2, t + h / 2, np.nan, w / h, h, 1.0]


@pytest.fixture(scope="module")
def shim_shares():
    d = os.path.join(HERE, "host_shim")
    so = os.path.join(d, "libshim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-x", "c++",
                           os.path.join(d, "shim.cpp"), "-o", so])
    L = C.CDLL(so)
    f32p = C.POINTER(C.c_float)
    L.shim_own_area_shares.argtypes = [f32p, C.c_int, f32p]

    def shares(b):
        b = np.ascontiguousarray(b, np.float32).reshape(-1, 6)
        out = np.zeros(len(b), np.float32)
        L.shim_own_area_shares(b.ctypes.data_as(f32p), len(b), out.ctypes.data_as(f32p))
        return out

    return shares


# ---- independent evaluation: inclusion-exclusion over the pieces box_i ∩ box_j, generic convex clipper
def _verts(b):
    xc, yc, a, asp, h = [float(v) for v in b[:5]]
    a = 0.0 if np.isnan(a) else a
    hw, hh = h * asp / 2, h / 2
    c, s = np.cos(a), np.sin(a)
    return [(xc + x * c - y * s, yc + x * s + y * c) for x, y in ((-hw, hh), (hw, hh), (hw, -hh), (-hw, -hh))]


def _sarea(p):
    return 0.5 * sum(p[i][0] * p[(i + 1) % len(p)][1] - p[(i + 1) % len(p)][0] * p[i][1] for i in range(len(p))) if len(p) >= 3 else 0.0


def _clip(subject, clipper):
    if _sarea(clipper) < 0:
        clipper = clipper[::-1]
    out = subject
    for i in range(len(clipper)):
        A, B = clipper[i], clipper[(i + 1) % len(clipper)]
        inp, out = out, []
        if not inp:
            break

        def side(P):
            return (B[0] - A[0]) * (P[1] - A[1]) - (B[1] - A[1]) * (P[0] - A[0])

        for k in range(len(inp)):
            P, Q = inp[k], inp[(k + 1) % len(inp)]
            sp, sq = side(P), side(Q)
            if sp >= 0:
                out.append(P)
            if (sp > 0 and sq < 0) or (sp < 0 and sq > 0):
                t = sp / (sp - sq)
                out.append((P[0] + t * (Q[0] - P[0]), P[1] + t * (Q[1] - P[1])))
    return out


def ie_share(boxes, i):
    vi = _verts(boxes[i])
    pieces = [c for c in (_clip(vi, _verts(b)) for j, b in enumerate(boxes) if j != i) if abs(_sarea(c)) > 0]
    tot = 0.0
    for r in range(1, len(pieces) + 1):
        for comb in itertools.combinations(range(len(pieces)), r):
            p = pieces[comb[0]]
            for q in comb[1:]:
                p = _clip(p, pieces[q])
                if abs(_sarea(p)) == 0:
                    break
            tot += (-1) ** (r + 1) * abs(_sarea(p))
    own = max(0.0, abs(_sarea(vi)) - tot)
    return min(1.0, own / (float(np.float32(boxes[i][3] * boxes[i][4] * boxes[i][4])) + 1e-5))


def random_boxes(r, n, oriented, span=100.0):
    b = np.zeros((n, 6), np.float32)
    b[:, 0] = r.uniform(0, span, n)
    b[:, 1] = r.uniform(0, span, n)
    b[:, 2] = r.uniform(-1.5, 1.5, n) if oriented else np.nan
    b[:, 3] = r.uniform(0.3, 0.8, n)
    b[:, 4] = r.uniform(40, 160, n)
    b[:, 5] = 1
    return b


def test_reference_known_answer(oracle):
    # bbox_own_areas.rs:50-79: 75 / 50 / 75 of 100 -> 0.75 / 0.50 / 0.75 within EPS
    got = oracle.own_area_shares([ltwh(0, 0, 10, 10), ltwh(5, 5, 10, 10), ltwh(10, 10, 10, 10)])
    assert np.all(np.abs(got - np.array([0.75, 0.5, 0.75], np.float32)) < 1e-5)


@pytest.mark.parametrize("oriented", [False, True])
def test_oracle_matches_inclusion_exclusion(oracle, oriented):
    r = np.random.default_rng(101 + oriented)
    worst = 0.0
    for trial in range(120):
        b = random_boxes(r, int(r.integers(2, 9)), oriented)
        got = oracle.own_area_shares(b)
        ref = np.array([ie_share(b.astype(np.float64), i) for i in range(len(b))])
        worst = max(worst, float(np.abs(got - ref).max()))
    assert worst < 2e-6, worst


def test_coincident_outlines(oracle):
    cases = {
        "identical": ([ltwh(0, 0, 10, 10), ltwh(0, 0, 10, 10)], [0, 0]),
        "triple duplicate": ([ltwh(0, 0, 10, 10)] * 3, [0, 0, 0]),
        "nested": ([ltwh(0, 0, 10, 10), ltwh(2, 2, 4, 4)], [0.84, 0]),
        "adjacent (shared edge, outside)": ([ltwh(0, 0, 10, 10), ltwh(10, 0, 10, 10)], [1, 1]),
        "duplicate cover of one half": ([ltwh(0, 0, 10, 10), ltwh(5, 0, 10, 10), ltwh(5, 0, 10, 10)], [0.5, 0, 0]),
        "two halves tile the box": ([ltwh(0, 0, 10, 10), ltwh(0, 0, 5, 10), ltwh(5, 0, 5, 10)], [0, 0, 0]),
        "cross": ([ltwh(0, 4, 10, 2), ltwh(4, 0, 2, 10)], [0.8, 0.8]),
        "far apart": ([ltwh(0, 0, 10, 10), ltwh(500, 500, 10, 10)], [1, 1]),
        "single": ([ltwh(0, 0, 10, 10)], [1]),
    }
    for name, (boxes, want) in cases.items():
        got = oracle.own_area_shares(boxes)
        assert np.all(np.abs(got - np.array(want, np.float32)) < 2e-5), (name, got)


@pytest.mark.parametrize("oriented", [False, True])
def test_product_arithmetic_matches_oracle(oracle, shim_shares, oriented):
    """sb_own_area.cuh (host build) == oracle: same formulation, independent code; also at scene scale (crowded)."""
    r = np.random.default_rng(55 + oriented)
    for trial in range(40):
        b = random_boxes(r, int(r.integers(1, 12)), oriented)
        np.testing.assert_allclose(shim_shares(b), oracle.own_area_shares(b), rtol=0, atol=1e-6)
    b = random_boxes(r, 300, oriented, span=1500.0)
    got, ref = shim_shares(b), oracle.own_area_shares(b)
    assert got.min() >= 0 and np.all(got != -1.0)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-6)
    assert 0.05 < float((ref < 0.999).mean()) < 1.0      # the scene does have overlaps


def test_oracle_tracker_computes_shares_when_thresholds_are_set(oracle):
    """visual_sort/simple_api.rs:110-127: with an own-area threshold the tracker derives the shares itself; passing the
    same shares explicitly must give identical results (and shares of 1 everywhere must not, in a crowded scene)."""
    r = np.random.default_rng(9)
    kw = dict(kind=2, positional_kind=1, iou_threshold=0.3, max_idle_epochs=3, visual_kind=0, visual_threshold=0.7,
              feature_dim=16, visual_max_observations=3, visual_min_votes=1, visual_minimal_track_length=1,
              visual_minimal_own_area_percentage_use=0.6, visual_minimal_own_area_percentage_collect=0.7)
    a, b, c = (oracle.Tracker(oracle.make_options(**kw)) for _ in range(3))
    base = random_boxes(r, 12, False, span=200.0)
    cent = r.standard_normal((12, 16)).astype(np.float32)
    differs = False
    for fr in range(5):
        boxes = base.copy()
        boxes[:, :2] += r.normal(0, 1.0, (12, 2)).astype(np.float32)
        feats = cent + 0.01 * r.standard_normal((12, 16)).astype(np.float32)
        own = oracle.own_area_shares(boxes)
        assert own.min() < 0.6 < own.max()                 # the threshold separates covered from free detections
        ra = a.predict_batch([0], [0, 12], boxes, features=feats)
        rb = b.predict_batch([0], [0, 12], boxes, features=feats, own_area=own)
        rc = c.predict_batch([0], [0, 12], boxes, features=feats, own_area=np.ones(12, np.float32))
        for key in ("ids", "epochs", "lengths", "voting_types"):
            assert np.array_equal(ra[key], rb[key]), (fr, key)
        fa, fb, fc = (t.scene_tracks(0)["feat_counts"] for t in (a, b, c))
        assert np.array_equal(fa, fb)
        differs = differs or not np.array_equal(fa, fc) or not np.array_equal(ra["voting_types"], rc["voting_types"])
    assert differs          # covered detections did not contribute features / visual votes
