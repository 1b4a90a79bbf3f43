"""sb_sincos.cuh -- the product's f64 sin / cos, one instruction sequence for host and device -- checked on the arguments a
box angle can take (f32 values widened to f64, src/utils/bbox.rs:287-330):

* it is CORRECTLY ROUNDED: every result equals mpmath's 200-bit value rounded to nearest (sampled), so the vertices of an
  oriented box are a pure function of the box on every platform and on the GPU;
* against this platform's C library (what the Rust reference calls here): identical bits for > 99 % of the angles; where
  they differ it is by one ulp and it is the C library that is not correctly rounded (glibc documents < 0.55 ulp).
  (CUDA's own sin / cos, used in round 1, are 1-2 ulp functions.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim():
    d = os.path.join(HERE, "host_shim")
    so = os.path.join(d, "libshim_sc.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-x", "c++",
                           os.path.join(d, "shim.cpp"), "-o", so])
    L = C.CDLL(so)
    for fn in (L.shim_sincos, L.shim_libm_sincos):
        fn.restype = None
        fn.argtypes = [C.POINTER(C.c_float), C.c_longlong, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return L


def both(L, a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = [np.empty(len(a), np.float64) for _ in range(4)]
    p = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    L.shim_sincos(a.ctypes.data_as(C.POINTER(C.c_float)), len(a), p(out[0]), p(out[1]))
    L.shim_libm_sincos(a.ctypes.data_as(C.POINTER(C.c_float)), len(a), p(out[2]), p(out[3]))
    return a, out


def check(L, angles, max_rate, n_verify=120):
    import mpmath as mp

    mp.mp.prec = 200
    a, (s, c, ls, lc) = both(L, angles)
    diff = (s.view(np.int64) != ls.view(np.int64)) | (c.view(np.int64) != lc.view(np.int64))
    rate = float(diff.mean()) if len(a) else 0.0
    assert rate <= max_rate, rate
    # differences from the C library are one ulp wide
    for ours, lib in ((s, ls), (c, lc)):
        d = np.abs(ours.view(np.int64) - lib.view(np.int64))
        assert d.max() <= 1
    # ours is the correctly rounded one: where the two differ (all of those, up to n_verify) and on a random sample
    idx = np.flatnonzero(diff)[:n_verify]
    rng = np.random.default_rng(1)
    idx = np.concatenate([idx, rng.integers(0, len(a), min(n_verify, len(a)))])
    for i in idx:
        x = mp.mpf(float(a[i]))
        assert float(mp.sin(x)) == s[i] and float(mp.cos(x)) == c[i], (float(a[i]), s[i], ls[i], c[i], lc[i])
    return rate


def test_ten_million_random_angles(shim):
    rng = np.random.default_rng(2024)
    worst = 0.0
    for lo, hi, n in ((-np.pi, np.pi, 6_000_000), (-1.6, 1.6, 2_000_000), (-100.0, 100.0, 1_500_000),
                      (-1e-3, 1e-3, 500_000)):
        worst = max(worst, check(shim, rng.uniform(lo, hi, n), 0.01))
    assert worst < 0.01


def test_neighbourhoods_of_multiples_of_half_pi_and_special_values(shim):
    """Arguments next to k * pi/2 (massive cancellation in the reduction), tiny, subnormal, zero, large."""
    vals = []
    for k in range(-2000, 2001):
        c = np.float32(k * np.pi / 2)
        x = c
        for _ in range(6):
            vals.append(x)
            x = np.nextafter(x, np.float32(np.inf), dtype=np.float32)
        x = c
        for _ in range(6):
            x = np.nextafter(x, np.float32(-np.inf), dtype=np.float32)
            vals.append(x)
    vals += [0.0, -0.0, 1e-45, -1e-45, 1e-30, 1.17549435e-38, 0.5, 1.0, 2.0, 3.0, 1e3, 12345.678, 1e5, 1.9e5]
    check(shim, np.array(vals, np.float32), 0.02, n_verify=400)
    # beyond the reduction's domain the library functions themselves are used: identical by definition
    a, (s, c, ls, lc) = both(shim, np.array([2.5e5, 1e9, 3e38], np.float32))
    assert np.array_equal(s, ls) and np.array_equal(c, lc)


def test_every_f32_in_a_dense_slab_around_typical_angles(shim):
    """All ~4 million consecutive f32 values of [0.5, 0.75): no gaps for a sampled test to miss."""
    lo = np.float32(0.5).view(np.uint32)
    bits = np.arange(lo, lo + (1 << 22), dtype=np.uint32)
    check(shim, bits.view(np.float32), 0.01)
