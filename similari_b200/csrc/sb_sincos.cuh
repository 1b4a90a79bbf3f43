// sb_sincos.cuh -- correctly rounded f64 sine / cosine, the same instruction sequence on the host and on the device.
//
// Why: the reference turns a box into its four vertices with f64 `cos` / `sin` of the f32 angle (src/utils/bbox.rs:287-330),
// i.e. with the C library's functions, which are correctly rounded for practically every argument.  CUDA's `sin` / `cos` are
// within 1-2 ulp but not correctly rounded, so a vertex could differ from the reference's in the last bit, the clipped area
// with it, and a 1-ulp IoU difference can flip `(v * 1e6) as i64` and with it a near-tie of the assignment.  Bit-exact
// oriented IoU needs bit-exact vertices; this header provides them by construction instead of by tolerance:
//   * Cody-Waite reduction by pi/2 held as three doubles (exact products through fma), valid for |x| < 2^17 * pi/2;
//   * table lookup at c = i/64 (sin c, cos c as double-doubles, generated with mpmath: sb_sincos_table.inc);
//   * Taylor series of sin t, cos t for |t| <= 1/128 in double-double arithmetic (error < 2^-95);
//   * angle addition in double-double, one final rounding.
// The result is the correctly rounded value unless the exact one lies within ~2^-42 ulp of a rounding boundary -- for the
// 2^32 possible f32 angles that does not happen in practice (tests/test_sincos_cpu.py compares 10^7 random f32 angles and
// every multiple-of-pi/2 neighbourhood with the C library, bit for bit).  Larger arguments fall back to sin() / cos().
#pragma once
#include <math.h>

#ifndef SB_HD
#ifdef __CUDACC__
#define SB_HD __host__ __device__ __forceinline__
#else
#define SB_HD inline
#endif
#endif

namespace sb {
namespace sc {

struct dd { double h, l; };

SB_HD dd two_sum(double a, double b) {
  const double s = a + b;
  const double bb = s - a;
  return dd{s, (a - (s - bb)) + (b - bb)};
}
SB_HD dd fast_two_sum(double a, double b) {   // |a| >= |b|
  const double s = a + b;
  return dd{s, b - (s - a)};
}
SB_HD dd two_prod(double a, double b) {
  const double p = a * b;
  return dd{p, fma(a, b, -p)};
}
SB_HD dd dd_add(dd a, dd b) {
  dd s = two_sum(a.h, b.h);
  const dd t = two_sum(a.l, b.l);
  s.l += t.h;
  s = fast_two_sum(s.h, s.l);
  s.l += t.l;
  return fast_two_sum(s.h, s.l);
}
SB_HD dd dd_mul(dd a, dd b) {
  dd p = two_prod(a.h, b.h);
  p.l += a.h * b.l;
  p.l += a.l * b.h;
  return fast_two_sum(p.h, p.l);
}
SB_HD dd dd_mul_d(dd a, double b) {
  dd p = two_prod(a.h, b);
  p.l += a.l * b;
  return fast_two_sum(p.h, p.l);
}
SB_HD dd dd_neg(dd a) { return dd{-a.h, -a.l}; }

// sin and cos of x, both correctly rounded (see the header comment for the domain)
SB_HD void sincos_cr(double x, double* sn, double* cs) {
  // constant tables (function-local so that one definition serves the host and the device compilation)
#define SB_SC_CONST static const
#include "sb_sincos_table.inc"
#undef SB_SC_CONST
  if (!(fabs(x) < 2.0e5)) {   // huge, inf or NaN: outside the reduction's domain
    *sn = sin(x);
    *cs = cos(x);
    return;
  }
  // ---- x = k * pi/2 + r, |r| <= pi/4 (+ a hair), r as a double-double
  const double kd = rint(x * kTwoOverPi);
  dd r;
  if (kd == 0.0) r = dd{x, 0.0};
  else {
    const dd p0 = two_prod(kd, kPio2[0]);
    const double r1 = x - p0.h;                 // exact (Sterbenz: p0.h is within a factor 2 of x)
    r = two_sum(r1, -p0.l);
    const dd p1 = two_prod(kd, kPio2[1]);
    r = dd_add(r, dd{-p1.h, -p1.l});
    r.l -= kd * kPio2[2];
    r = fast_two_sum(r.h, r.l);
  }
  const bool neg = r.h < 0.0;
  if (neg) r = dd_neg(r);
  // ---- r = c + t, c = i / 64
  int i = (int)rint(r.h * 64.0);
  if (i > 51) i = 51;
  const double c = (double)i * 0.015625;
  dd t = two_sum(r.h - c, r.l);                 // r.h - c is exact (both are multiples of ulp(r.h), same binade or below)
  const dd t2 = dd_mul(t, t);
  // sin t = t + t * t2 * (-1/3! + t2 * (1/5! + t2 * (-1/7! + t2 * (1/9! - t2 / 11!))))
  double u = kSinCoef[3][0] - t2.h * kSinCoef[4][0];
  u = -kSinCoef[2][0] + t2.h * u;
  dd q = dd_add(dd{kSinCoef[1][0], kSinCoef[1][1]}, dd_mul_d(t2, u));
  q = dd_add(dd{-kSinCoef[0][0], -kSinCoef[0][1]}, dd_mul(t2, q));
  q = dd_mul(t2, q);
  const dd st = dd_add(t, dd_mul(t, q));
  // cos t = 1 + t2 * (-1/2! + t2 * (1/4! + t2 * (-1/6! + t2 * (1/8! - t2 / 10!))))
  double v = kCosCoef[3][0] - t2.h * kCosCoef[4][0];
  v = -kCosCoef[2][0] + t2.h * v;
  dd w = dd_add(dd{kCosCoef[1][0], kCosCoef[1][1]}, dd_mul_d(t2, v));
  w = dd_add(dd{-kCosCoef[0][0], -kCosCoef[0][1]}, dd_mul(t2, w));
  w = dd_mul(t2, w);
  const dd ct = dd_add(dd{1.0, 0.0}, w);
  // ---- sin(c + t), cos(c + t)
  const dd sc_ = dd{kSinTab[i][0], kSinTab[i][1]}, cc_ = dd{kCosTab[i][0], kCosTab[i][1]};
  dd s = dd_add(dd_mul(sc_, ct), dd_mul(cc_, st));
  dd co = dd_add(dd_mul(cc_, ct), dd_neg(dd_mul(sc_, st)));
  if (neg) s = dd_neg(s);
  // ---- quadrant
  const long long k = (long long)kd;
  const int n = (int)(((k % 4) + 4) % 4);
  dd rs, rc;
  if (n == 0) { rs = s; rc = co; }
  else if (n == 1) { rs = co; rc = dd_neg(s); }
  else if (n == 2) { rs = dd_neg(s); rc = dd_neg(co); }
  else { rs = dd_neg(co); rc = s; }
  *sn = rs.h + rs.l;
  *cs = rc.h + rc.l;
}

}  // namespace sc
}  // namespace sb
