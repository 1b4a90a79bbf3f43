// Host build of the product's per-pair arithmetic (similari_b200/csrc/sb_math.cuh) so that the CPU test-suite can
// compare it bit-for-bit with the oracle before any GPU time is spent.  TEST INFRASTRUCTURE: the product never
// loads this library and has no CPU execution path.
#include <string.h>
#include "../../similari_b200/csrc/sb_math.cuh"
#include "../../similari_b200/csrc/sb_own_area.cuh"
extern "C" {
void shim_vertices(const float* b, double* out8) { sb::box_vertices(b[0], b[1], b[2], b[3], b[4], out8); }
double shim_clip_area(const double* s8, const double* c8) { return sb::clip_area(s8, c8); }
float shim_iou(const float* l, const float* r) {
  if (sb::too_far(l[0], l[1], sb::box_radius(l[3], l[4]), r[0], r[1], sb::box_radius(r[3], r[4]))) return nanf("");
  double vl[8], vr[8];
  sb::box_vertices(l[0], l[1], l[2], l[3], l[4], vl);
  sb::box_vertices(r[0], r[1], r[2], r[3], r[4], vr);
  return sb::iou_from_area(sb::clip_area(vl, vr), l[4], l[3], r[4], r[3]);
}
void shim_kalman_initiate(float pw, float vw, const float* b, float* st) {
  sb::Box bx{b[0], b[1], b[2], b[3], b[4], b[5]};
  sb::kalman_initiate(pw, vw, bx, st);
}
void shim_kalman_predict(float pw, float vw, const float* in, float* out) { sb::kalman_predict(pw, vw, in, out); }
void shim_kalman_update(float pw, const float* in, const float* b, float* out) {
  sb::Box bx{b[0], b[1], b[2], b[3], b[4], b[5]};
  sb::kalman_update(pw, in, bx, out);
}
float shim_maha(float pw, const float* st, const float* b) {
  float l5[5];
  for (int i = 0; i < 5; ++i) l5[i] = sqrtf(sb::kalman_proj_var(pw, st[4], st[10 + 4 * i], i));
  return sb::maha_distance(st, l5, b[0], b[1], sb::angle_or0(b[2]), b[3], b[4]);
}
long long shim_weight(float v) { return sb::weight_i64(v); }
void shim_own_area_shares(const float* boxes, int n, float* out) {
  for (int i = 0; i < n; ++i) out[i] = sb::own_area_share_seq(boxes, n, i);
}
double shim_overlap_bound(const double* a8, const double* b8) { return sb::rect_overlap_bound(a8, b8); }
int shim_iou_bound_fails(const float* l, const float* r, float conf, float thr) {
  return sb::iou_bound_fails(l[4], l[3], r[4], r[3], conf, thr) ? 1 : 0;
}
}
// sincos_cr (sb_sincos.cuh) and the C library's sin / cos over an array of f32 angles
extern "C" void shim_sincos(const float* angles, long long n, double* s, double* c) {
  for (long long i = 0; i < n; ++i) sb::sc::sincos_cr((double)angles[i], &s[i], &c[i]);
}
extern "C" void shim_libm_sincos(const float* angles, long long n, double* s, double* c) {
  for (long long i = 0; i < n; ++i) { s[i] = sin((double)angles[i]); c[i] = cos((double)angles[i]); }
}
