// ops.cu -- stateless operators of the C ABI (sb200_sort_cost_matrix, sb200_visual_cost_matrix, sb200_sort_voting,
// sb200_visual_voting, sb200_kalman_*).  They drive the SAME kernels as the tracker's predict path on a one-scene
// scratch store, so a parity test of an operator is a parity test of the product kernel.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/similari_b200.h"
#include "sb_engine.cuh"

extern "C" void sb200__set_error(const char* msg);  // engine.cu

namespace {

struct Scratch {
  std::vector<void*> ptrs;
  cudaStream_t st = nullptr;
  ~Scratch() {
    for (void* p : ptrs) cudaFree(p);
    if (st) cudaStreamDestroy(st);
  }
  template <typename T>
  T* alloc(size_t n, bool zero = false) {
    void* p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(1, n) * sizeof(T)) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    if (zero) cudaMemsetAsync(p, 0, std::max<size_t>(1, n) * sizeof(T), st);
    return reinterpret_cast<T*>(p);
  }
  template <typename T>
  T* upload(const T* h, size_t n) {
    T* d = alloc<T>(n);
    if (d && n) cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, st);
    return d;
  }
};

int ops_fail(int code, const std::string& msg) {
  sb200__set_error(msg.c_str());
  return code;
}
int begin(Scratch& sc, int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return ops_fail(SB200_ERR_CUDA, "no CUDA device available (this library has no CPU execution path)");
  }
  if (device < 0 || device >= n) return ops_fail(SB200_ERR_INVALID, "device out of range");
  if (cudaSetDevice(device) != cudaSuccess) return ops_fail(SB200_ERR_CUDA, "cudaSetDevice failed");
  if (cudaStreamCreateWithFlags(&sc.st, cudaStreamNonBlocking) != cudaSuccess) return ops_fail(SB200_ERR_CUDA, "cudaStreamCreate failed");
  return 0;
}
int finish(Scratch& sc) {
  cudaError_t e = cudaStreamSynchronize(sc.st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return ops_fail(SB200_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e));
  return 0;
}

sb::Params base_params() {
  sb::Params p;
  memset(&p, 0, sizeof(p));
  p.max_idle_epochs = 1;
  p.max_obs = 1;
  p.d8 = 8;
  return p;
}

__global__ void track_geom_kernel(int iou, const float* boxes, int n, float* radius, double* vert, unsigned int* epoch) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* b = boxes + (size_t)i * 6;
  radius[i] = sb::box_radius(b[3], b[4]);
  epoch[i] = 0;
  if (iou) sb::box_vertices(b[0], b[1], b[2], b[3], b[4], vert + (size_t)i * 8);
}
__global__ void fill_u8_kernel(unsigned char* p, unsigned char v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
}  // namespace

extern "C" {

int sb200_sort_cost_matrix(int32_t positional_kind, float iou_threshold, float min_confidence, float pos_weight,
                           float vel_weight, const float* cand_boxes, int32_t m, const float* track_boxes,
                           const float* track_states30, int32_t n, float* out_mn, int32_t device) {
  if (m < 0 || n < 0 || (m > 0 && !cand_boxes) || (n > 0 && !track_boxes) || (m > 0 && n > 0 && !out_mn))
    return ops_fail(SB200_ERR_INVALID, "bad arguments");
  if (positional_kind == SB200_POS_MAHA && n > 0 && !track_states30)
    return ops_fail(SB200_ERR_INVALID, "track_states30 is required for the Mahalanobis metric");
  Scratch sc;
  int rc = begin(sc, device);
  if (rc) return rc;
  if (m == 0 || n == 0) return 0;
  sb::Params p = base_params();
  p.positional_kind = positional_kind;
  p.iou_threshold = iou_threshold;
  p.min_confidence = min_confidence;
  p.pos_weight = pos_weight;
  p.vel_weight = vel_weight;
  sb::TrackStore ts;
  memset(&ts, 0, sizeof(ts));
  ts.track_cap = n;
  ts.pred = sc.upload(track_boxes, (size_t)n * 6);
  ts.radius = sc.alloc<float>(n);
  ts.epoch = sc.alloc<unsigned int>(n);
  ts.vert = sc.alloc<double>((size_t)n * 8);
  ts.kst_stride = 30;
  ts.kst = track_states30 ? sc.upload(track_states30, (size_t)n * 30) : sc.alloc<float>((size_t)n * 30, true);
  sb::Frame f;
  memset(&f, 0, sizeof(f));
  f.total = m;
  f.in_boxes = sc.upload(cand_boxes, (size_t)m * 6);
  f.c_box = sc.alloc<float>((size_t)m * 6);
  f.c_radius = sc.alloc<float>(m);
  f.c_conf = sc.alloc<float>(m);
  f.c_vert = sc.alloc<double>((size_t)m * 8);
  f.pos = sc.alloc<float>((size_t)m * n);
  f.pos_total = (long long)m * n;
  f.pos_dense_all = true;               // the operator returns the dense matrix
  f.pos_cnt = sc.alloc<int>(4, true);   // sparse list disabled here (capacity 0): only the dense matrix is returned
  f.pos_list = sc.alloc<sb::PosEntry>(1);
  sb::SceneDesc d;
  memset(&d, 0, sizeof(d));
  d.m = m; d.n = n; d.epoch = 1;
  f.scenes = sc.upload(&d, 1);
  if (!ts.pred || !ts.radius || !ts.epoch || !ts.vert || !ts.kst || !f.in_boxes || !f.c_box || !f.c_radius ||
      !f.c_conf || !f.c_vert || !f.pos || !f.scenes)
    return ops_fail(SB200_ERR_CUDA, "cudaMalloc failed");
  track_geom_kernel<<<(n + 127) / 128, 128, 0, sc.st>>>(positional_kind == SB200_POS_IOU, ts.pred, n, ts.radius, ts.vert, ts.epoch);
  sb::launch_prep(p, f, 1, m, sc.st);
  sb::launch_pos_cost(p, ts, f, 1, m, n, sc.st);
  cudaMemcpyAsync(out_mn, f.pos, (size_t)m * n * 4, cudaMemcpyDeviceToHost, sc.st);
  return finish(sc);
}

int sb200_visual_cost_matrix(int32_t visual_kind, float threshold, const float* cand_features, int32_t m,
                             const float* track_features, int32_t n, int32_t d, float* out_mn, int32_t device) {
  if (m < 0 || n < 0 || d <= 0 || (m > 0 && !cand_features) || (n > 0 && !track_features) || (m > 0 && n > 0 && !out_mn))
    return ops_fail(SB200_ERR_INVALID, "bad arguments");
  Scratch sc;
  int rc = begin(sc, device);
  if (rc) return rc;
  if (m == 0 || n == 0) return 0;
  sb::Params p = base_params();
  p.is_visual = true;
  p.visual_kind = visual_kind;
  p.visual_threshold = threshold;
  p.feature_dim = d;
  p.d8 = (d + 7) / 8 * 8;
  p.max_obs = 1;
  p.min_track_length = 0;
  // track side: norms through the candidate-norm kernel on a scratch frame
  sb::Frame ft;
  memset(&ft, 0, sizeof(ft));
  ft.total = n;
  ft.in_feat = sc.upload(track_features, (size_t)n * d);
  ft.in_boxes = sc.alloc<float>((size_t)n * 6, true);
  ft.c_box = sc.alloc<float>((size_t)n * 6);
  ft.c_radius = sc.alloc<float>(n);
  ft.c_conf = sc.alloc<float>(n);
  ft.c_flags = sc.alloc<unsigned char>(n);
  ft.c_norm2 = sc.alloc<float>(n, true);
  sb::TrackStore ts;
  memset(&ts, 0, sizeof(ts));
  ts.track_cap = n;
  ts.pred = sc.alloc<float>((size_t)n * 6, true);
  ts.radius = sc.alloc<float>(n, true);
  ts.epoch = sc.alloc<unsigned int>(n, true);
  ts.feat = sc.alloc<float>((size_t)n * p.d8, true);
  ts.obs_phys = sc.alloc<unsigned char>(n, true);
  ts.obs_hasf = sc.alloc<unsigned char>(n);
  ts.obs_n = sc.alloc<unsigned char>(n);
  ts.feat_cnt = sc.alloc<unsigned char>(n);
  sb::Frame f;
  memset(&f, 0, sizeof(f));
  f.total = m;
  f.in_feat = sc.upload(cand_features, (size_t)m * d);
  f.in_boxes = sc.alloc<float>((size_t)m * 6, true);
  f.c_box = sc.alloc<float>((size_t)m * 6);
  f.c_radius = sc.alloc<float>(m);
  f.c_conf = sc.alloc<float>(m);
  f.c_flags = sc.alloc<unsigned char>(m);
  f.c_norm2 = sc.alloc<float>(m, true);
  f.vis = sc.alloc<float>((size_t)m * n);
  sb::SceneDesc sd;
  memset(&sd, 0, sizeof(sd));
  sd.m = m; sd.n = n; sd.nb = n; sd.epoch = 1;   // stateless operator: one observation per track, block == track
  f.scenes = sc.upload(&sd, 1);
  if (!ft.in_feat || !ft.in_boxes || !ft.c_box || !ft.c_radius || !ft.c_conf || !ft.c_flags || !ft.c_norm2 || !ts.pred ||
      !ts.radius || !ts.epoch || !ts.feat || !ts.obs_phys || !ts.obs_hasf || !ts.obs_n || !ts.feat_cnt || !f.in_feat ||
      !f.in_boxes || !f.c_box || !f.c_radius || !f.c_conf || !f.c_flags || !f.c_norm2 || !f.vis || !f.scenes)
    return ops_fail(SB200_ERR_CUDA, "cudaMalloc failed");
  sb::launch_prep(p, ft, 1, n, sc.st);
  ts.fnorm2 = ft.c_norm2;
  cudaMemcpy2DAsync(ts.feat, (size_t)p.d8 * 4, ft.in_feat, (size_t)d * 4, (size_t)d * 4, n, cudaMemcpyDeviceToDevice, sc.st);
  fill_u8_kernel<<<(n + 255) / 256, 256, 0, sc.st>>>(ts.obs_hasf, 1, n);
  fill_u8_kernel<<<(n + 255) / 256, 256, 0, sc.st>>>(ts.obs_n, 1, n);
  fill_u8_kernel<<<(n + 255) / 256, 256, 0, sc.st>>>(ts.feat_cnt, 1, n);
  sb::launch_prep(p, f, 1, m, sc.st);
  fill_u8_kernel<<<(m + 255) / 256, 256, 0, sc.st>>>(f.c_flags, 3, m);
  // same kernel selection rule as the tracker (engine.cu): tensor-core screen + exact refinement for large
  // contractions with a selective threshold; SB200_VIS_KERNEL=simt|tc overrides
  sb::TcArgs tc;
  memset(&tc, 0, sizeof(tc));
  const bool selective = visual_kind == SB200_VIS_EUCLIDEAN ? (threshold < 1e18f) : (threshold > -1.0f);
  tc.use_tc = selective && p.d8 >= 64 && (long long)m * n * p.d8 >= (1ll << 28);
  if (const char* ev = getenv("SB200_VIS_KERNEL")) {
    if (!strcmp(ev, "simt")) tc.use_tc = false;
    else if (!strcmp(ev, "tc")) tc.use_tc = true;
  }
  f.scene_max = sc.alloc<unsigned int>(1);
  cudaDeviceGetAttribute(&tc.num_sms, cudaDevAttrMultiProcessorCount, device);
  {
    int lcap = std::max(4096, m * 64);
    if (const char* ev = getenv("SB200_VIS_PAIR_CAP")) lcap = std::max(1, atoi(ev));
    sd.vis_lbase = 0; sd.vis_lcap = lcap; sd.pos_lbase = 0; sd.pos_lcap = 0;
    cudaMemcpyAsync(f.scenes, &sd, sizeof(sd), cudaMemcpyHostToDevice, sc.st);
    f.vis_pairs = sc.alloc<sb::VisPair>((size_t)lcap);
    f.vis_val = sc.alloc<float>((size_t)lcap);
    f.pos_cnt = sc.alloc<int>(8, true);
    f.vis_cnt = f.pos_cnt + 1;
    f.scene_mode = f.pos_cnt + 2;
    f.vis_mode = f.pos_cnt + 3;
    f.refine_next = f.pos_cnt + 4;
    f.dense_cnt = f.pos_cnt + 5;
    f.pos_list = sc.alloc<sb::PosEntry>(1);
    if (!f.vis_pairs || !f.vis_val || !f.pos_cnt || !f.pos_list) return ops_fail(SB200_ERR_CUDA, "cudaMalloc failed");
  }
  if (tc.use_tc) {
    std::vector<sb::TcTile> tiles;
    {
        // SB200_SCREEN = single | multicast | pair (default): CTA organisation of the screen kernel
        const char* e = getenv("SB200_SCREEN");
        tc.cluster2 = !(e && !strcmp(e, "single")) && getenv("SB200_SCREEN_SINGLE") == nullptr;
        tc.pair = tc.cluster2 && !(e && !strcmp(e, "multicast"));
      }
    for (int m0 = 0; m0 < m; m0 += (tc.cluster2 ? 256 : 128))
      for (int c0 = 0; c0 < n; c0 += 256) tiles.push_back(sb::TcTile{0, m0, c0, 0});
    tc.n_tiles = (int)tiles.size();
    tc.d_tiles = sc.upload(tiles.data(), tiles.size());
    tc.a_rows = m;
    tc.b_rows = n;
    f.c_bf16 = sc.alloc<unsigned short>((size_t)m * p.d8);
    ts.feat_bf16 = sc.alloc<unsigned short>((size_t)n * p.d8);
    tc.colmeta = sc.alloc<sb::VisColMeta>(n + 256);   // the screen kernel bulk-copies whole 256-column slabs
    tc.colgeo = sc.alloc<sb::VisColGeo>(n);
    tc.colb = sc.alloc<float>(n + 256);
    tc.colvalid = sc.alloc<unsigned int>((n + 256) / 32 + 4);
    tc.rowmeta = sc.alloc<sb::VisRowMeta>(m + 256);
    tc.total_cols = n;
    if (!tc.d_tiles || !f.c_bf16 || !ts.feat_bf16 || !tc.colmeta || !tc.colgeo || !tc.rowmeta || !tc.colb || !tc.colvalid)
      return ops_fail(SB200_ERR_CUDA, "cudaMalloc failed");
    sb::launch_to_bf16(ft.in_feat, d, d, p.d8, n, ts.feat_bf16, sc.st);
    sb::launch_to_bf16(f.in_feat, d, d, p.d8, m, f.c_bf16, sc.st);   // (the tracker fuses this into cand_norm_kernel)
  }
  ts.fnorm2 = ft.c_norm2;
  int vr = sb::launch_vis_cost(p, ts, f, 1, m, n, tc, sc.st);
  if (vr == 0 && tc.use_tc) sb::launch_vis_densify(p, f, 1, sc.st);
  if (vr != 0) return ops_fail(SB200_ERR_CUDA, "visual cost launch failed");
  cudaMemcpyAsync(out_mn, f.vis, (size_t)m * n * 4, cudaMemcpyDeviceToHost, sc.st);
  return finish(sc);
}

static int run_voting(bool visual, float threshold, int min_votes, const float* pos_mn, const float* vis_mnk, int m, int n,
                      int k, int32_t* winner, uint8_t* voting_type, int device) {
  if (m < 0 || n < 0 || (m > 0 && !winner) || (m > 0 && n > 0 && !pos_mn)) return ops_fail(SB200_ERR_INVALID, "bad arguments");
  if (visual && (k < 1 || k > sb::kMaxObs || (m > 0 && n > 0 && !vis_mnk))) return ops_fail(SB200_ERR_INVALID, "bad arguments");
  Scratch sc;
  int rc = begin(sc, device);
  if (rc) return rc;
  if (m == 0) return 0;
  sb::Params p = base_params();
  p.positional_kind = SB200_POS_IOU;  // threshold is taken verbatim: (threshold * 1e6) as i64
  p.iou_threshold = threshold;
  p.is_visual = visual;
  p.max_obs = visual ? k : 1;
  p.min_votes = min_votes;
  sb::TrackStore ts;
  memset(&ts, 0, sizeof(ts));
  sb::Frame f;
  memset(&f, 0, sizeof(f));
  f.total = m;
  f.pos = sc.upload(pos_mn, (size_t)m * n);
  if (visual) f.vis = sc.upload(vis_mnk, (size_t)m * n * k);
  f.winner = sc.alloc<int>(m);
  f.c_vt = sc.alloc<unsigned char>(m);
  f.new_count = sc.alloc<int>(1);
  {
    // operators take dense matrices: scene mode 1 routes the request to the dense voting kernel
    int hm[4] = {0, 0, 1, 0};
    f.pos_cnt = sc.upload(hm, 4);
    if (!f.pos_cnt) return ops_fail(SB200_ERR_CUDA, "cudaMalloc failed");
    f.vis_cnt = f.pos_cnt + 1;
    f.scene_mode = f.pos_cnt + 2;
    f.vis_mode = f.pos_cnt + 3;
  }
  sb::SceneDesc sd;
  memset(&sd, 0, sizeof(sd));
  sd.m = m; sd.n = n; sd.epoch = 1;
  f.scenes = sc.upload(&sd, 1);
  if (!f.pos || (visual && !f.vis) || !f.winner || !f.c_vt || !f.new_count || !f.scenes) return ops_fail(SB200_ERR_CUDA, "cudaMalloc failed");
  if (visual) {
    f.scene_max = sc.alloc<unsigned int>(1);
    if (!f.scene_max) return ops_fail(SB200_ERR_CUDA, "cudaMalloc failed");
    sb::launch_scene_max(p, f, 1, /*init_only=*/true, sc.st);
    if (n > 0) sb::launch_scene_max(p, f, 1, /*init_only=*/false, sc.st);
  }
  int vr = sb::launch_voting(p, ts, f, 1, m, n, sc.st);
  if (vr == -3) return ops_fail(SB200_ERR_CAPACITY, "scene too large for the on-chip assignment solver");
  if (vr != 0) return ops_fail(SB200_ERR_CUDA, std::string("voting launch failed: ") + cudaGetErrorString((cudaError_t)vr));
  cudaMemcpyAsync(winner, f.winner, (size_t)m * 4, cudaMemcpyDeviceToHost, sc.st);
  if (voting_type) cudaMemcpyAsync(voting_type, f.c_vt, (size_t)m, cudaMemcpyDeviceToHost, sc.st);
  return finish(sc);
}

int sb200_sort_voting(float threshold, const float* cost_mn, int32_t m, int32_t n, int32_t* winner, int32_t device) {
  return run_voting(false, threshold, 0, cost_mn, nullptr, m, n, 1, winner, nullptr, device);
}

int sb200_visual_voting(float positional_threshold, int32_t min_votes, const float* pos_mn, const float* vis_mnk,
                        int32_t m, int32_t n, int32_t k, int32_t* winner, uint8_t* voting_type, int32_t device) {
  return run_voting(true, positional_threshold, min_votes, pos_mn, vis_mnk, m, n, k, winner, voting_type, device);
}

int sb200_own_area_shares(const float* boxes, int32_t n, float* out, int32_t device) {
  if (n < 0 || (n > 0 && (!boxes || !out))) return ops_fail(SB200_ERR_INVALID, "bad arguments");
  Scratch sc;
  int rc = begin(sc, device);
  if (rc) return rc;
  if (n == 0) return 0;
  sb::Frame f;
  memset(&f, 0, sizeof(f));
  f.total = n;
  sb::SceneDesc sd;
  memset(&sd, 0, sizeof(sd));
  sd.m = n;
  f.scenes = sc.upload(&sd, 1);
  f.status = sc.alloc<int>(1, true);
  float* d_boxes = sc.upload(boxes, (size_t)n * 6);
  float* d_out = sc.alloc<float>(n);
  int* d_ovf_cnt = sc.alloc<int>(1);
  int2* d_ovf = sc.alloc<int2>(n);
  if (!f.scenes || !f.status || !d_boxes || !d_out || !d_ovf_cnt || !d_ovf) return ops_fail(SB200_ERR_CUDA, "cudaMalloc failed");
  sb::launch_own_area(f, 1, n, d_boxes, d_out, d_ovf_cnt, d_ovf, sc.st);
  int status = 0;
  cudaMemcpyAsync(out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, sc.st);
  cudaMemcpyAsync(&status, f.status, sizeof(int), cudaMemcpyDeviceToHost, sc.st);
  rc = finish(sc);
  if (rc) return rc;
  if (status & 2) return ops_fail(SB200_ERR_CAPACITY, "more than 2800 boxes overlap one box (own-area shares)");
  return 0;
}

static int kalman_op(int op, float pw, float vw, const float* in30, const float* boxes, int n, float* out30, int device) {
  if (n < 0 || (n > 0 && !out30) || (op != 0 && n > 0 && !in30) || (op != 1 && n > 0 && !boxes)) return ops_fail(SB200_ERR_INVALID, "bad arguments");
  Scratch sc;
  int rc = begin(sc, device);
  if (rc) return rc;
  if (n == 0) return 0;
  float* din = in30 ? sc.upload(in30, (size_t)n * 30) : nullptr;
  float* db = boxes ? sc.upload(boxes, (size_t)n * 6) : nullptr;
  float* dout = sc.alloc<float>((size_t)n * 30);
  if ((in30 && !din) || (boxes && !db) || !dout) return ops_fail(SB200_ERR_CUDA, "cudaMalloc failed");
  sb::launch_kalman_ops(op, pw, vw, din, db, n, dout, sc.st);
  cudaMemcpyAsync(out30, dout, (size_t)n * 120, cudaMemcpyDeviceToHost, sc.st);
  return finish(sc);
}
int sb200_kalman_initiate(float pw, float vw, const float* boxes, int32_t n, float* states30, int32_t device) {
  return kalman_op(0, pw, vw, nullptr, boxes, n, states30, device);
}
int sb200_kalman_predict(float pw, float vw, const float* in30, int32_t n, float* out30, int32_t device) {
  return kalman_op(1, pw, vw, in30, nullptr, n, out30, device);
}
int sb200_kalman_update(float pw, float vw, const float* in30, const float* boxes, int32_t n, float* out30, int32_t device) {
  return kalman_op(2, pw, vw, in30, boxes, n, out30, device);
}

}  // extern "C"
