"""CPU check of the product's device arithmetic (similari_b200/csrc/sb_math.cuh, host-compiled by tests/host_shim)
against the oracle: bit-exact Kalman block forms, bit-exact clipped area, IoU.  No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim():
    d = os.path.join(HERE, "host_shim")
    so = os.path.join(d, "libshim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-x", "c++",
                           os.path.join(d, "shim.cpp"), "-o", so])
    L = C.CDLL(so)
    f32p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
    L.shim_vertices.argtypes = [f32p, f64p]
    L.shim_clip_area.argtypes = [f64p, f64p]
    L.shim_clip_area.restype = C.c_double
    L.shim_iou.argtypes = [f32p, f32p]
    L.shim_iou.restype = C.c_float
    L.shim_kalman_initiate.argtypes = [C.c_float, C.c_float, f32p, f32p]
    L.shim_kalman_predict.argtypes = [C.c_float, C.c_float, f32p, f32p]
    L.shim_kalman_update.argtypes = [C.c_float, f32p, f32p, f32p]
    L.shim_maha.argtypes = [C.c_float, f32p, f32p]
    L.shim_maha.restype = C.c_float
    L.shim_weight.argtypes = [C.c_float]
    L.shim_weight.restype = C.c_longlong
    L.shim_overlap_bound.argtypes = [f64p, f64p]
    L.shim_overlap_bound.restype = C.c_double
    L.shim_iou_bound_fails.argtypes = [f32p, f32p, C.c_float, C.c_float]
    L.shim_iou_bound_fails.restype = C.c_int
    return L


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def pack30(st110):
    """oracle mean[10]+cov[10x10] -> product mean[10] + 5 x (Pii, Pi,i+5, Pi+5,i, Pi+5,i+5)."""
    mean, cov = st110[:10], st110[10:].reshape(10, 10)
    out = np.zeros(30, np.float32)
    out[:10] = mean
    for i in range(5):
        out[10 + 4 * i: 14 + 4 * i] = [cov[i, i], cov[i, i + 5], cov[i + 5, i], cov[i + 5, i + 5]]
    return out


def offblock_zero(st110):
    cov = st110[10:].reshape(10, 10).copy()
    for i in range(5):
        cov[i, i] = cov[i, i + 5] = cov[i + 5, i] = cov[i + 5, i + 5] = 0
    return not cov.any()


def rand_box(rng, oriented):
    return np.array([rng.uniform(0, 1920), rng.uniform(0, 1080),
                     rng.uniform(-1.5, 1.5) if oriented else np.nan,
                     rng.uniform(0.3, 0.8), rng.uniform(40, 160), rng.uniform(0.3, 1.0)], np.float32)


@pytest.mark.parametrize("oriented", [False, True])
def test_kalman_block_forms_bit_exact(oracle, shim, oriented):
    rng = np.random.default_rng(7 + oriented)
    pw, vw = np.float32(1 / 20), np.float32(1 / 160)
    for trial in range(50):
        b = rand_box(rng, oriented)
        ref = oracle.kalman_initiate(b, pw, vw)
        mine = np.zeros(30, np.float32)
        shim.shim_kalman_initiate(pw, vw, fp(b), fp(mine))
        assert np.array_equal(pack30(ref), mine)
        for step in range(12):
            ref = oracle.kalman_predict(ref, pw, vw)
            nxt = np.zeros(30, np.float32)
            shim.shim_kalman_predict(pw, vw, fp(mine), fp(nxt))
            mine = nxt
            assert offblock_zero(ref)
            assert np.array_equal(pack30(ref), mine), (trial, step, "predict")
            z = b.copy()
            z[:2] += rng.normal(0, 2, 2).astype(np.float32)
            z[3:5] *= rng.uniform(0.98, 1.02, 2).astype(np.float32)
            if oriented:
                z[2] += np.float32(rng.normal(0, 0.02))
            # Mahalanobis distance of the measurement against the *current* state
            d_ref = oracle.kalman_distance(ref, z, pw, vw)
            d_mine = shim.shim_maha(pw, fp(mine), fp(z))
            assert np.float32(d_ref) == np.float32(d_mine), (trial, step, d_ref, d_mine)
            ref = oracle.kalman_update(ref, z, pw, vw)
            nxt = np.zeros(30, np.float32)
            shim.shim_kalman_update(pw, fp(mine), fp(z), fp(nxt))
            mine = nxt
            assert offblock_zero(ref)
            assert np.array_equal(pack30(ref), mine), (trial, step, "update")
            b = z


@pytest.mark.parametrize("oriented", [False, True])
def test_clip_area_and_iou_bit_exact(oracle, shim, oriented):
    rng = np.random.default_rng(11 + oriented)
    n_some = 0
    for trial in range(3000):
        l = rand_box(rng, oriented)
        r = l.copy()
        r[:2] += rng.normal(0, 30, 2).astype(np.float32)
        r[3:5] *= rng.uniform(0.7, 1.3, 2).astype(np.float32)
        if oriented:
            r[2] += np.float32(rng.normal(0, 0.5))
        vl, vr = np.zeros(8), np.zeros(8)
        shim.shim_vertices(fp(l), dp(vl))
        shim.shim_vertices(fp(r), dp(vr))
        if not oriented:
            assert np.array_equal(vl.reshape(4, 2), oracle.vertices(l))
        else:
            # the product's sin / cos are correctly rounded (sb_sincos.cuh, tests/test_sincos_cpu.py); the oracle calls the
            # C library like the reference does, which is off by one ulp for ~0.3 % of the angles: vertices agree to that
            np.testing.assert_allclose(vl.reshape(4, 2), oracle.vertices(l), rtol=0, atol=2e-13)
        # the clip itself, on identical vertices: bit for bit
        ol, orr = np.ascontiguousarray(oracle.vertices(l), np.float64).ravel(), np.ascontiguousarray(oracle.vertices(r), np.float64).ravel()
        a_ref = oracle.polygon_area(oracle.sh_clip(oracle.vertices(l), oracle.vertices(r)))
        a_mine = shim.shim_clip_area(dp(ol), dp(orr))
        assert a_ref == a_mine, (trial, a_ref, a_mine)
        i_ref = oracle.iou(l, r)
        i_mine = shim.shim_iou(fp(l), fp(r))
        if i_ref is None:
            assert np.isnan(i_mine) or (oriented and i_mine < 1e-9)
        else:
            n_some += 1
            if not oriented:
                assert np.float32(i_ref) == np.float32(i_mine)
            else:
                assert abs(float(np.float32(i_ref)) - float(np.float32(i_mine))) <= 1.2e-7 * max(float(i_ref), 1e-3)
    assert n_some > 500


def test_identical_and_touching_boxes(oracle, shim):
    b = oracle.ltwh(0.0, 0.0, 3.0, 5.0)
    assert shim.shim_iou(fp(b), fp(b)) == np.float32(oracle.iou(b, b))
    c = oracle.ltwh(3.0, 0.0, 3.0, 5.0)  # shares an edge: sliver area from the f32 aspect rounding, or None
    ref, mine = oracle.iou(b, c), shim.shim_iou(fp(b), fp(c))
    assert (ref is None and np.isnan(mine)) or np.float32(ref) == np.float32(mine)
    d = oracle.ltwh(4.0, 0.0, 2.0, 4.0)  # exactly representable, disjoint but within 2R => clip area exactly 0
    assert oracle.iou(oracle.ltwh(0.0, 0.0, 2.0, 4.0), d) is None
    assert np.isnan(shim.shim_iou(fp(oracle.ltwh(0.0, 0.0, 2.0, 4.0)), fp(d)))


def test_weight_cast(oracle, shim):
    for v in [0.6, 0.69, 0.3, 100.0 / 0.05, 0.0, float("nan"), 1e30, -1e30]:
        w = shim.shim_weight(np.float32(v))
        if np.isnan(v):
            assert w == 0
        elif abs(v) > 1e20:
            assert w in (2**63 - 1, -2**63)
        else:
            assert w == int(np.float32(v) * np.float32(1e6))


@pytest.mark.parametrize("oriented", [False, True])
def test_iou_pregates_never_reject_a_pair_the_reference_keeps(shim, oriented):
    """The culled positional kernel skips the f64 clip when a pre-gate proves the reference's result is None
    (rect_overlap_bound == 0: separated; IoU upper bounds below the threshold).  Properties on 40k random pairs,
    incl. near-duplicates, touching and nested boxes: the bound is never below the clipped area, a zero bound means
    an (at most rounding-sliver) empty clip, and no pair whose exact IoU * conf reaches the threshold is rejected."""
    r = np.random.default_rng(77 + oriented)
    n = 40000
    thr = 0.3
    rejected = kept = 0
    for i in range(n):
        l = np.array([r.uniform(0, 400), r.uniform(0, 400), r.uniform(-1.6, 1.6) if oriented else np.nan,
                      r.uniform(0.3, 0.8), r.uniform(40, 160), 1.0], np.float32)
        mode = i % 4
        if mode == 0:      # near duplicate
            q = l.copy()
            q[:2] += r.normal(0, 3, 2).astype(np.float32)
            q[4] *= np.float32(r.uniform(0.9, 1.1))
            if oriented:
                q[2] += np.float32(r.normal(0, 0.05))
        elif mode == 1:    # neighbour
            q = np.array([l[0] + r.uniform(-150, 150), l[1] + r.uniform(-150, 150),
                          r.uniform(-1.6, 1.6) if oriented else np.nan, r.uniform(0.3, 0.8), r.uniform(40, 160), 1.0], np.float32)
        elif mode == 2:    # exactly touching / shifted by its own width
            q = l.copy()
            q[0] += l[3] * l[4]
        else:              # nested
            q = l.copy()
            q[4] *= np.float32(0.5)
        conf = np.float32(r.uniform(0.1, 1.0))
        vl, vq = np.zeros(8), np.zeros(8)
        shim.shim_vertices(fp(l), dp(vl))
        shim.shim_vertices(fp(q), dp(vq))
        area = shim.shim_clip_area(dp(vl), dp(vq))
        ub = shim.shim_overlap_bound(dp(vl), dp(vq))
        s_ = float(np.float32(l[4] * l[4] * l[3] + q[4] * q[4] * q[3]))
        assert ub >= area * (1 - 1e-12) or (ub == 0.0 and area <= 1e-6 * s_), (i, ub, area)
        iou = area / (s_ - area) if area > 0 else 0.0
        passes = (np.float32(iou) * conf) >= np.float32(thr) and area != 0.0
        gate = ub == 0.0 or (ub < 0.5 * s_ and ub * float(conf) * 1.0001 < thr * (s_ - ub)) or \
            bool(shim.shim_iou_bound_fails(fp(l), fp(q), conf, np.float32(thr)))
        assert not (gate and passes), (i, iou, conf, ub, area)
        rejected += gate
        kept += (not gate)
    assert rejected > n // 4 and kept > n // 8     # both branches are exercised
