// kernels_nms.cu -- greedy non-maximum suppression of (oriented) boxes, src/utils/nms.rs:32-72.
//
//   filter  : score.unwrap_or(MAX) > score_threshold.unwrap_or(f32::MIN) && height > 0 && aspect > 0
//   rank    : score.unwrap_or(height), stable sort descending
//   suppress: a kept box cb removes every later ob with (intersection(cb, ob) as f32) / ob.area() > nms_threshold
//
// GPU shape: (1) rank by counting (stable by construction, O(n^2) compares, no library sort);
// (2) 64x64-tiled suppression bit-mask with the f64 Sutherland-Hodgman clip of sb_math.cuh behind the
// circumscribed-circle gate; (3) one warp sweeps the mask rows in rank order, the `removed` bitmap lives in
// registers and the rows are prefetched (row loads do not depend on the keep/drop decision).
#include <cstring>
#include <string>

#include "../../include/similari_b200.h"
#include "sb_engine.cuh"

extern "C" void sb200__set_error(const char* msg);

namespace sb {

__global__ void nms_filter_kernel(const float* boxes, const float* scores, int n, float score_thr, float* rank,
                                  unsigned char* valid) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* b = boxes + (size_t)i * 6;
  bool has = scores != nullptr && !is_nan(scores[i]);
  float s = has ? scores[i] : 3.402823466e+38f;
  bool ok = s > score_thr && b[4] > 0.0f && b[3] > 0.0f;
  valid[i] = ok;
  rank[i] = has ? scores[i] : b[4];
}

// position of box i in the stable descending order of the valid boxes
__global__ void nms_rank_kernel(const float* rank, const unsigned char* valid, int n, int* order, int* n_valid) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float s_rank[256];
  __shared__ unsigned char s_valid[256];
  const bool mine = i < n && valid[i];
  const float r = i < n ? rank[i] : 0.0f;
  int pos = 0;
  for (int base = 0; base < n; base += 256) {
    int j = base + threadIdx.x;
    s_rank[threadIdx.x] = j < n ? rank[j] : 0.0f;
    s_valid[threadIdx.x] = j < n ? valid[j] : 0;
    __syncthreads();
    if (mine) {
      int lim = min(256, n - base);
      for (int q = 0; q < lim; ++q) {
        if (!s_valid[q]) continue;
        float rq = s_rank[q];
        int jq = base + q;
        if (rq > r || (rq == r && jq < i)) ++pos;
      }
    }
    __syncthreads();
  }
  if (mine) { order[pos] = i; atomicAdd(n_valid, 1); }
}

__global__ void nms_geom_kernel(const float* boxes, const int* order, const int* n_valid, float* sx, float* sy,
                                float* sr, float* sarea, double* vert) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= *n_valid) return;
  const float* b = boxes + (size_t)order[k] * 6;
  sx[k] = b[0]; sy[k] = b[1];
  sr[k] = box_radius(b[3], b[4]);
  sarea[k] = box_area(b[3], b[4]);
  box_vertices(b[0], b[1], b[2], b[3], b[4], vert + (size_t)k * 8);
}

// mask[i][jb] bit t: box i suppresses box jb*64+t (only j > i)
__global__ void __launch_bounds__(64) nms_mask_kernel(const int* n_valid, float nms_thr, const float* sx, const float* sy,
                                                      const float* sr, const float* sarea, const double* vert,
                                                      unsigned long long* mask, int words) {
  const int nv = *n_valid;
  const int ib = blockIdx.y, jb = blockIdx.x;
  if (jb < ib || ib * 64 >= nv || jb * 64 >= nv) return;
  __shared__ float cx[64], cy[64], cr[64], ca[64];
  __shared__ double cv[64][8];
  const int t = threadIdx.x;
  const int j = jb * 64 + t;
  if (j < nv) {
    cx[t] = sx[j]; cy[t] = sy[j]; cr[t] = sr[j]; ca[t] = sarea[j];
    for (int q = 0; q < 8; ++q) cv[t][q] = vert[(size_t)j * 8 + q];
  }
  __syncthreads();
  const int i = ib * 64 + t;
  if (i >= nv) return;
  const float ix = sx[i], iy = sy[i], ir = sr[i];
  double iv[8];
  for (int q = 0; q < 8; ++q) iv[q] = vert[(size_t)i * 8 + q];
  unsigned long long bits = 0;
  const int lim = min(64, nv - jb * 64);
  for (int q = 0; q < lim; ++q) {
    const int jj = jb * 64 + q;
    if (jj <= i) continue;
    // intersection(cb, ob): 0.0 behind the circumscribed-circle gate, else clip(subject = cb, clip = ob)
    const double a = too_far(ix, iy, ir, cx[q], cy[q], cr[q]) ? 0.0 : clip_area(iv, cv[q]);
    const float metric = (float)a / ca[q];
    if (metric > nms_thr) bits |= 1ull << q;
  }
  mask[(size_t)i * words + jb] = bits;
}

// one warp; lane l owns words l, l+32, ... of the removed bitmap (<= 16 words per lane => 32768 boxes per pass;
// larger inputs loop over word groups)
__global__ void __launch_bounds__(32) nms_sweep_kernel(const int* n_valid, const unsigned long long* mask, int words,
                                                       const int* order, int* out_idx, int* out_count) {
  const int nv = *n_valid;
  const int lane = threadIdx.x;
  extern __shared__ unsigned long long removed[];  // words entries
  for (int w = lane; w < words; w += 32) removed[w] = 0ull;
  __syncwarp();
  int kept = 0;
  for (int i = 0; i < nv; ++i) {
    const unsigned long long rw = removed[i >> 6];
    if ((rw >> (i & 63)) & 1ull) continue;
    if (lane == 0) out_idx[kept] = order[i];
    ++kept;
    // rows only carry bits for columns >= the row's own block
    const unsigned long long* row = mask + (size_t)i * words;
    for (int w = (i >> 6) + lane; w < words; w += 32) removed[w] |= row[w];
    __syncwarp();
  }
  if (lane == 0) *out_count = kept;
}

int launch_nms(const float* d_boxes, const float* d_scores, int n, float nms_thr, float score_thr, int has_score_thr,
               int* d_out_idx, int* d_out_count, cudaStream_t st) {
  if (n == 0) return 0;
  const float sthr = has_score_thr ? score_thr : -3.402823466e+38f;  // f32::MIN
  float *rank, *sx, *sy, *sr, *sa;
  unsigned char* valid;
  int *order, *nvalid;
  double* vert;
  unsigned long long* mask;
  const int words = (n + 63) / 64;
  cudaError_t e;
#define NA(p, bytes) if ((e = cudaMalloc(&p, (bytes))) != cudaSuccess) return (int)e;
  NA(rank, 4 * (size_t)n) NA(sx, 4 * (size_t)n) NA(sy, 4 * (size_t)n) NA(sr, 4 * (size_t)n) NA(sa, 4 * (size_t)n)
  NA(valid, (size_t)n) NA(order, 4 * (size_t)n) NA(nvalid, 4) NA(vert, 64 * (size_t)n)
  NA(mask, 8 * (size_t)n * words)
#undef NA
  cudaMemsetAsync(nvalid, 0, 4, st);
  cudaMemsetAsync(mask, 0, 8 * (size_t)n * words, st);
  nms_filter_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_boxes, d_scores, n, sthr, rank, valid);
  nms_rank_kernel<<<(n + 255) / 256, 256, 0, st>>>(rank, valid, n, order, nvalid);
  nms_geom_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_boxes, order, nvalid, sx, sy, sr, sa, vert);
  dim3 grid(words, words);
  nms_mask_kernel<<<grid, 64, 0, st>>>(nvalid, nms_thr, sx, sy, sr, sa, vert, mask, words);
  size_t smem = 8 * (size_t)words;
  if (smem > 48 * 1024) cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  nms_sweep_kernel<<<1, 32, smem, st>>>(nvalid, mask, words, order, d_out_idx, d_out_count);
  note_launch(5);
  e = cudaStreamSynchronize(st);
  cudaFree(rank); cudaFree(sx); cudaFree(sy); cudaFree(sr); cudaFree(sa); cudaFree(valid); cudaFree(order);
  cudaFree(nvalid); cudaFree(vert); cudaFree(mask);
  return (int)e;
}

}  // namespace sb

extern "C" int64_t sb200_nms(const float* boxes, const float* scores, int32_t n, float nms_threshold, float score_threshold,
                             int32_t has_score_threshold, int32_t* out_idx, int32_t device) {
  auto fail = [](int code, const std::string& m) { sb200__set_error(m.c_str()); return (int64_t)code; };
  if (n < 0 || (n > 0 && (!boxes || !out_idx))) return fail(SB200_ERR_INVALID, "bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return fail(SB200_ERR_CUDA, "no CUDA device available (this library has no CPU execution path)"); }
  if (device < 0 || device >= ndev) return fail(SB200_ERR_INVALID, "device out of range");
  if (n == 0) return 0;
  if ((size_t)((n + 63) / 64) * 8 > 200 * 1024) return fail(SB200_ERR_CAPACITY, "nms: too many boxes for the on-chip sweep");
  cudaSetDevice(device);
  cudaStream_t st;
  if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return fail(SB200_ERR_CUDA, "cudaStreamCreate failed");
  float *db = nullptr, *ds = nullptr;
  int *dout = nullptr, *dcnt = nullptr;
  int64_t ret = 0;
  cudaError_t e = cudaMalloc(&db, 24 * (size_t)n);
  if (e == cudaSuccess && scores) e = cudaMalloc(&ds, 4 * (size_t)n);
  if (e == cudaSuccess) e = cudaMalloc(&dout, 4 * (size_t)n);
  if (e == cudaSuccess) e = cudaMalloc(&dcnt, 4);
  if (e == cudaSuccess) {
    cudaMemcpyAsync(db, boxes, 24 * (size_t)n, cudaMemcpyHostToDevice, st);
    if (scores) cudaMemcpyAsync(ds, scores, 4 * (size_t)n, cudaMemcpyHostToDevice, st);
    int rc = sb::launch_nms(db, ds, n, nms_threshold, score_threshold, has_score_threshold, dout, dcnt, st);
    if (rc != 0) e = (cudaError_t)rc;
  }
  int cnt = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&cnt, dcnt, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess && cnt > 0) e = cudaMemcpy(out_idx, dout, 4 * (size_t)cnt, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) ret = fail(SB200_ERR_CUDA, std::string("nms: CUDA error: ") + cudaGetErrorString(e));
  else ret = cnt;
  cudaFree(db); cudaFree(ds); cudaFree(dout); cudaFree(dcnt);
  cudaStreamDestroy(st);
  return ret;
}
