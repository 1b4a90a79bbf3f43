"""The oracle's feature distances exist twice: scalar lane-by-lane loops (the restatement proper) and 8-lane AVX2
vectors (the speed of a SIMD build of the reference, used for the timed CPU baseline).  Both must agree bit for bit --
every GPU parity test compares the device's f32 values with whatever path the host picks."""
import ctypes as C

import numpy as np


def _fns(oracle):
    L = oracle.lib()
    f32p = C.POINTER(C.c_float)
    for name in ("orc_euclidean_scalar", "orc_cosine_scalar", "orc_euclidean_blocks", "orc_cosine_blocks"):
        getattr(L, name).argtypes = [f32p, f32p, C.c_int]
        getattr(L, name).restype = C.c_float
    L.orc_simd_active.restype = C.c_int
    return L, f32p


def test_simd_and_scalar_feature_distances_are_bit_identical(oracle):
    L, f32p = _fns(oracle)
    r = np.random.default_rng(2024)
    n_checked = 0
    for blocks in (1, 2, 3, 8, 16, 64, 65, 256):
        for scale in (1e-3, 1.0, 37.0, 1e4):
            for _ in range(40):
                a = (r.standard_normal(blocks * 8) * scale).astype(np.float32)
                b = (r.standard_normal(blocks * 8) * scale).astype(np.float32)
                if _ % 7 == 0:            # unit vectors around a common centroid, like the ReID workload
                    a /= np.linalg.norm(a)
                    b = a + (0.02 * r.standard_normal(blocks * 8)).astype(np.float32)
                    b /= np.linalg.norm(b)
                if _ % 11 == 0:           # zero padding tail (Feature::from_vec)
                    a[-5:] = 0
                    b[-5:] = 0
                pa, pb = a.ctypes.data_as(f32p), b.ctypes.data_as(f32p)
                for sc, fast in ((L.orc_euclidean_scalar, L.orc_euclidean_blocks), (L.orc_cosine_scalar, L.orc_cosine_blocks)):
                    x, y = np.float32(sc(pa, pb, blocks)), np.float32(fast(pa, pb, blocks))
                    assert x.tobytes() == y.tobytes() or (np.isnan(x) and np.isnan(y)), (blocks, scale, x, y)
                    n_checked += 1
    assert n_checked > 2000


def test_simd_path_reports_itself(oracle):
    L, _ = _fns(oracle)
    assert L.orc_simd_active() in (0, 1)      # 1 on AVX2 hosts unless ORACLE_NO_SIMD is set
