// kernels_cost.cu -- candidate preparation and the positional cost matrix (IoU / Mahalanobis).
//
// Replaces, for every (candidate, track) pair of a scene, the reference chain
//   TrackStore::foreign_track_distances -> Track::distances (compatible gate, src/track.rs:604-652)
//   -> SortMetric::metric / VisualMetric::positional_metric (src/trackers/sort/metric.rs:38-77,
//      src/trackers/visual_sort/metric.rs:156-198)
// Roofline: HBM-write bound -- 4 B of cost per pair-association out, (m + n) small per-box records in.
#include "sb_engine.cuh"

namespace sb {

// --------------------------------------------------------------------------------------------------------
// prep: one thread per detection. Candidate track construction (sort/simple_api.rs:125-145): the Kalman
// initiate->predict->update of a fresh state leaves the box unchanged except angle == 0.0 -> None
// (src/utils/kalman.rs:82-86); confidence is preserved (kalman_prediction.rs:28-29).
__global__ void prep_kernel(Params p, Frame f) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= f.total) return;
  const float* b = f.in_boxes + (size_t)i * 6;
  float xc = b[0], yc = b[1], ang = b[2], asp = b[3], h = b[4], conf = b[5];
  if (ang == 0.0f) ang = nanf("");
  float* cb = f.c_box + (size_t)i * 6;
  cb[0] = xc; cb[1] = yc; cb[2] = ang; cb[3] = asp; cb[4] = h; cb[5] = conf;
  f.c_radius[i] = box_radius(asp, h);
  f.c_conf[i] = conf < p.min_confidence ? p.min_confidence : conf;
  if (p.positional_kind == 1) box_vertices(xc, yc, ang, asp, h, f.c_vert + (size_t)i * 8);
  if (p.is_visual) {
    bool hasf = f.in_feat != nullptr && (f.in_hasf == nullptr || f.in_hasf[i] != 0);
    float q = f.in_quality ? f.in_quality[i] : 1.0f;
    // VisualMetric::feature_can_be_used with the *_use thresholds, visual_sort/metric.rs:227-249
    bool ok = q >= p.min_quality_use;
    if (p.use_own_area && f.in_own) ok = ok && (f.in_own[i] >= p.min_own_use);
    ok = ok && (box_area(asp, h) >= p.min_area);
    f.c_flags[i] = (unsigned char)((hasf ? 1 : 0) | ((hasf && ok) ? 2 : 0));
  }
}

// squared norms of candidate features in the reference's order (cosine, src/distance.rs:36-44):
// per 8-lane block reduce_add, blocks accumulated sequentially. One warp per detection.
__device__ __forceinline__ float reduce_add8(const float* t) {
  float q0 = t[0] + t[4], q1 = t[1] + t[5], q2 = t[2] + t[6], q3 = t[3] + t[7];
  float d0 = q0 + q2, d1 = q1 + q3;
  return d0 + d1;
}

__global__ void cand_norm_kernel(Params p, Frame f) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= f.total) return;
  const int nblk = p.d8 / 8;
  const float* row = f.in_feat + (size_t)w * p.feature_dim;
  float acc = 0.0f;
  // blocks are accumulated in order; lanes compute block sums in parallel, lane 0 folds them sequentially
  for (int base = 0; base < nblk; base += 32) {
    int blk = base + lane;
    float bs = 0.0f;
    if (blk < nblk) {
      float t[8];
#pragma unroll
      for (int l = 0; l < 8; ++l) {
        int d = blk * 8 + l;
        float v = d < p.feature_dim ? row[d] : 0.0f;
        t[l] = v * v;
      }
      bs = reduce_add8(t);
    }
    int cnt = min(32, nblk - base);
    for (int j = 0; j < cnt; ++j) {
      float v = __shfl_sync(0xffffffffu, bs, j);
      acc = acc + v;
    }
  }
  if (lane == 0) f.c_norm2[w] = acc;
}

void launch_prep(const Params& p, const Frame& f, int n_scenes, int max_m, cudaStream_t st) {
  (void)n_scenes; (void)max_m;
  if (f.total == 0) return;
  prep_kernel<<<(f.total + 255) / 256, 256, 0, st>>>(p, f);
  if (p.is_visual && f.in_feat) {  // squared norms (cosine; euclidean on the tensor-core path)
    long long threads = (long long)f.total * 32;
    cand_norm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(p, f);
  }
}

// --------------------------------------------------------------------------------------------------------
// positional cost: tile = TM candidates x TN tracks per CTA of TN threads*TY; thread owns one track column
// (its record lives in registers) and walks TM/TY candidates staged in shared memory (broadcast reads).
constexpr int TN = 128;
constexpr int TM = 16;
constexpr int TY = 2;

struct CandTile {
  float xc[TM], yc[TM], ang0[TM], asp[TM], h[TM], r[TM], conf[TM];
  double vert[TM][8];
};

template <int POS>
__global__ void __launch_bounds__(TN * TY) pos_cost_kernel(Params p, TrackStore ts, Frame f) {
  const SceneDesc sc = f.scenes[blockIdx.z];
  const int n0 = blockIdx.x * TN, m0 = blockIdx.y * TM;
  if (n0 >= sc.n || m0 >= sc.m) return;
  __shared__ CandTile ct;
  const int tid = threadIdx.y * TN + threadIdx.x;
  // stage candidates
  for (int i = tid; i < TM; i += TN * TY) {
    int m = m0 + i;
    if (m < sc.m) {
      int g = sc.det_base + m;
      const float* cb = f.c_box + (size_t)g * 6;
      ct.xc[i] = cb[0]; ct.yc[i] = cb[1]; ct.ang0[i] = angle_or0(cb[2]); ct.asp[i] = cb[3]; ct.h[i] = cb[4];
      ct.r[i] = f.c_radius[g]; ct.conf[i] = f.c_conf[g];
    }
  }
  if (POS == 1) {
    for (int i = tid; i < TM * 8; i += TN * TY) {
      int m = m0 + i / 8;
      if (m < sc.m) ct.vert[i / 8][i % 8] = f.c_vert[(size_t)(sc.det_base + m) * 8 + (i % 8)];
    }
  }
  __syncthreads();
  const int n = n0 + threadIdx.x;
  if (n >= sc.n) return;
  const size_t ti = (size_t)sc.slot * ts.track_cap + n;
  const float* tb = ts.pred + ti * 6;
  const float txc = tb[0], tyc = tb[1], tasp = tb[3], th = tb[4];
  const float tr = ts.radius[ti];
  const unsigned int tep = ts.epoch[ti];
  float mean5[5], l5[5];
  double tv[8];
  if (POS == 0) {
    const float* st = ts.kst + ti * kStateFloats;
    const float hh = st[4];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      mean5[i] = st[i];
      l5[i] = sqrtf(kalman_proj_var(p.pos_weight, hh, st[10 + 4 * i], i));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) tv[i] = ts.vert[ti * 8 + i];
  }
  float* out = f.pos + sc.pos_off;
  const float qnan = nanf("");
#pragma unroll 1
  for (int i = threadIdx.y; i < TM; i += TY) {
    int m = m0 + i;
    if (m >= sc.m) break;
    float v = qnan;
    const float cx = ct.xc[i], cy = ct.yc[i], cr = ct.r[i];
    if (compat_ok(p, sc.epoch, tep, cx, cy, cr, txc, tyc, tr) && !too_far(cx, cy, cr, txc, tyc, tr)) {
      if (POS == 0) {
        float d = maha_distance(mean5, l5, cx, cy, ct.ang0[i], ct.asp[i], ct.h[i]);
        v = maha_cost(d) / ct.conf[i];
      } else {
        double a = clip_area(ct.vert[i], tv);
        float iou = iou_from_area(a, ct.h[i], ct.asp[i], th, tasp);
        if (!is_nan(iou)) {
          iou = iou * ct.conf[i];
          v = iou >= p.iou_threshold ? iou : qnan;
        }
      }
    }
    out[(size_t)m * sc.n + n] = v;
    // sparse view for the voting stage: valid entries are rare (gated by 2R and the threshold), append them
    const bool valid = !is_nan(v);
    const unsigned am = __activemask();
    const unsigned bal = __ballot_sync(am, valid);
    if (bal) {
      const int lane = threadIdx.x & 31;
      const int leader = __ffs(bal) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&f.pos_cnt[blockIdx.z], __popc(bal));
      base = __shfl_sync(am, base, leader);
      if (valid) {
        const int slot = base + __popc(bal & ((1u << lane) - 1));
        if (slot < sc.pos_lcap) {
          PosEntry e; e.m = (unsigned short)m; e.n = (unsigned short)n; e.v = v;
          f.pos_list[sc.pos_lbase + slot] = e;
        }
      }
    }
  }
}

void launch_pos_cost(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_m, int max_n,
                     cudaStream_t st) {
  if (n_scenes == 0 || max_m == 0 || max_n == 0) return;
  dim3 grid((max_n + TN - 1) / TN, (max_m + TM - 1) / TM, n_scenes);
  dim3 block(TN, TY);
  if (p.positional_kind == 0) pos_cost_kernel<0><<<grid, block, 0, st>>>(p, ts, f);
  else pos_cost_kernel<1><<<grid, block, 0, st>>>(p, ts, f);
}

// --------------------------------------------------------------------------------------------------------
// visual distances, fp32 SIMT in the reference's summation order (src/distance.rs:9-47): per 8-lane block a
// horizontal reduce_add, blocks accumulated sequentially.  One CTA computes a VM x VN tile of
// (candidate, track-observation) pairs; operands are staged through shared memory in chunks of VK floats.
// Output layout vis[m][n][k] (k = logical observation index of the track), NaN == None after the threshold /
// gate logic of VisualMetric::metric (src/trackers/visual_sort/metric.rs:200-225,253-295).
constexpr int VM = 64, VN = 64, VK = 32, VT = 256;  // 4x4 pairs per thread

__global__ void __launch_bounds__(VT) vis_cost_kernel(Params p, TrackStore ts, Frame f) {
  if (f.scene_mode[blockIdx.z] == 0) return;  // this scene's visual entries come from the screen + refine path
  const SceneDesc sc = f.scenes[blockIdx.z];
  const int K = p.max_obs;
  const int ncols = sc.n * K;  // column c = n*K + k (logical obs)
  const int c0 = blockIdx.x * VN, m0 = blockIdx.y * VM;
  if (c0 >= ncols || m0 >= sc.m) return;
  __shared__ float sa[VM][VK + 1];
  __shared__ float sb_[VN][VK + 1];
  __shared__ int col_row[VN];      // feature row (idx*K + phys) or -1
  __shared__ float col_norm[VN];
  __shared__ unsigned char col_ok[VN];
  __shared__ unsigned char row_ok[VM];
  __shared__ float row_norm[VM];
  const int tid = threadIdx.x;
  const bool cosine = p.visual_kind == 1;
  if (tid < VN) {
    int c = c0 + tid;
    int row = -1; float nrm = 0.0f; unsigned char ok = 0;
    if (c < ncols) {
      int n = c / K, k = c % K;
      size_t ti = (size_t)sc.slot * ts.track_cap + n;
      if (k < ts.obs_n[ti] && ts.obs_hasf[ti * K + k] && ts.feat_cnt[ti] >= p.min_track_length) {
        int phys = ts.obs_phys[ti * K + k];
        row = (int)(ti * K + phys);
        nrm = cosine ? ts.fnorm2[ti * K + phys] : 0.0f;
        ok = 1;
      }
    }
    col_row[tid] = row; col_norm[tid] = nrm; col_ok[tid] = ok;
  }
  if (tid >= 64 && tid < 64 + VM) {
    int i = tid - 64;
    int m = m0 + i;
    unsigned char ok = 0; float nrm = 0.0f;
    if (m < sc.m) {
      int g = sc.det_base + m;
      ok = (f.c_flags[g] & 2) ? 1 : 0;
      nrm = cosine ? f.c_norm2[g] : 0.0f;
    }
    row_ok[i] = ok; row_norm[i] = nrm;
  }
  __syncthreads();
  const int tx = tid % 16, ty = tid / 16;  // thread computes rows ty*4..+3, cols tx*4..+3
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
  const int D = p.feature_dim, D8 = p.d8;
  for (int k0 = 0; k0 < D8; k0 += VK) {
    // stage VM x VK candidate floats and VN x VK track floats (zero padded)
    for (int e = tid; e < VM * VK; e += VT) {
      int r = e / VK, c = e % VK;
      int m = m0 + r, d = k0 + c;
      float v = 0.0f;
      if (m < sc.m && d < D && row_ok[r]) v = f.in_feat[(size_t)(sc.det_base + m) * D + d];
      sa[r][c] = v;
    }
    for (int e = tid; e < VN * VK; e += VT) {
      int r = e / VK, c = e % VK;
      int d = k0 + c;
      float v = 0.0f;
      int row = col_row[r];
      if (row >= 0 && d < D8) v = ts.feat[(size_t)row * D8 + d];
      sb_[r][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int blk = 0; blk < VK / 8; ++blk) {
      float av[4][8], bv[4][8];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int l = 0; l < 8; ++l) av[a][l] = sa[ty * 4 + a][blk * 8 + l];
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int l = 0; l < 8; ++l) bv[b][l] = sb_[tx * 4 + b][blk * 8 + l];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          float t[8];
          if (cosine) {
#pragma unroll
            for (int l = 0; l < 8; ++l) t[l] = av[a][l] * bv[b][l];
          } else {
#pragma unroll
            for (int l = 0; l < 8; ++l) { float df = av[a][l] - bv[b][l]; t[l] = df * df; }
          }
          acc[a][b] = acc[a][b] + reduce_add8(t);
        }
    }
    __syncthreads();
  }
  float* out = f.vis + sc.vis_off;
  const float qnan = nanf("");
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int r = ty * 4 + a, m = m0 + r;
    if (m >= sc.m) continue;
    int g = sc.det_base + m;
    const float cx = f.c_box[(size_t)g * 6], cy = f.c_box[(size_t)g * 6 + 1], cr = f.c_radius[g];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int cc = tx * 4 + b, c = c0 + cc;
      if (c >= ncols) continue;
      float v = qnan;
      if (row_ok[r] && col_ok[cc]) {
        int n = c / K;
        size_t ti = (size_t)sc.slot * ts.track_cap + n;
        const float* tb = ts.pred + ti * 6;
        if (compat_ok(p, sc.epoch, ts.epoch[ti], cx, cy, cr, tb[0], tb[1], ts.radius[ti])) {
          if (cosine) {
            float d = acc[a][b] / sqrtf(row_norm[r] * col_norm[cc]);
            if (d >= p.visual_threshold) v = 1.0f - d;   // is_ok + distance_to_weight
          } else {
            float d = sqrtf(acc[a][b]);
            if (d <= p.visual_threshold) v = d;
          }
        }
      }
      out[(size_t)m * ncols + c] = v;
    }
  }
}

int launch_vis_cost(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_m, int max_n,
                    const TcArgs& tc, cudaStream_t st) {
  if (n_scenes == 0) return 0;
  if (!p.is_visual) {
    launch_scene_mode(p, f, n_scenes, false, st);
    return 0;
  }
  const bool any = max_m > 0 && max_n > 0 && f.in_feat != nullptr;
  const bool use_tc = tc.use_tc && any;
  launch_scene_max(p, f, n_scenes, /*init_only=*/true, st);
  if (use_tc) {
    // tensor-core screen -> per-scene survivor lists
    launch_to_bf16(f.in_feat, p.feature_dim, p.feature_dim, p.d8, f.total, f.c_bf16, st);
    int rc = launch_vis_cost_tc(p, ts, f, n_scenes, max_n, tc, /*phase=*/0, st);
    if (rc != 0) return rc;
  }
  launch_scene_mode(p, f, n_scenes, use_tc, st);   // which scenes stay sparse, which fall back to the dense kernels
  if (use_tc) {
    int rc = launch_vis_cost_tc(p, ts, f, n_scenes, max_n, tc, /*phase=*/1, st);   // exact refinement of the survivors
    if (rc != 0) return rc;
  }
  if (max_m > 0 && max_n > 0) {
    dim3 grid((max_n * p.max_obs + VN - 1) / VN, (max_m + VM - 1) / VM, n_scenes);
    vis_cost_kernel<<<grid, VT, 0, st>>>(p, ts, f);
    launch_scene_max(p, f, n_scenes, /*init_only=*/false, st);
  }
  return 0;
}

}  // namespace sb
