// engine.cu -- host side of libsimilari_b200.so: the C ABI of include/similari_b200.h, device memory management
// and the per-frame launch sequence (prep -> positional cost -> visual cost -> voting -> apply).
// There is no CPU execution path in this library: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/similari_b200.h"
#include "sb_engine.cuh"

#include <atomic>

namespace sb {
static std::atomic<unsigned long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
}  // namespace sb

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess)                                                                     \
      return fail(SB200_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// grow-only device buffer
struct DBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    size_t nb = std::max(need, bytes + bytes / 2);
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, nb);
    if (e != cudaSuccess) return fail(SB200_ERR_CUDA, "cudaMalloc(%zu) failed: %s", nb, cudaGetErrorString(e));
    if (p) cudaFree(p);
    p = np;
    bytes = nb;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HBuf {  // grow-only pinned host buffer, mapped into the device address space (dp)
  void* p = nullptr;
  void* dp = nullptr;   // device-side alias: kernels can read the buffer over PCIe without a copy-engine transfer
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    size_t nb = std::max(need, bytes + bytes / 2);
    void* np = nullptr;
    cudaError_t e = cudaHostAlloc(&np, nb, cudaHostAllocMapped);
    if (e != cudaSuccess) return fail(SB200_ERR_CUDA, "cudaHostAlloc(%zu) failed: %s", nb, cudaGetErrorString(e));
    void* ndp = nullptr;
    e = cudaHostGetDevicePointer(&ndp, np, 0);
    if (e != cudaSuccess) { cudaFreeHost(np); return fail(SB200_ERR_CUDA, "cudaHostGetDevicePointer failed: %s", cudaGetErrorString(e)); }
    if (p) cudaFreeHost(p);
    p = np;
    dp = ndp;
    bytes = nb;
    return 0;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    dp = nullptr;
    bytes = 0;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

int make_params(const sb200_options& o, sb::Params* out) {
  sb::Params p;
  memset(&p, 0, sizeof(p));
  if (o.kind < 0 || o.kind > 3) return fail(SB200_ERR_INVALID, "unknown tracker kind %d", o.kind);
  if (o.positional_kind != SB200_POS_MAHA && o.positional_kind != SB200_POS_IOU)
    return fail(SB200_ERR_INVALID, "unknown positional metric %d", o.positional_kind);
  p.kind = o.kind;
  p.positional_kind = o.positional_kind;
  p.visual_kind = o.visual_kind;
  p.iou_threshold = o.iou_threshold;
  p.min_confidence = o.min_confidence;
  p.pos_weight = o.kalman_position_weight;
  p.vel_weight = o.kalman_velocity_weight;
  p.max_idle_epochs = o.max_idle_epochs;
  if (o.max_idle_epochs < 0) return fail(SB200_ERR_INVALID, "max_idle_epochs must be >= 0");
  if (o.n_constraints < 0 || o.n_constraints > SB200_MAX_CONSTRAINTS)
    return fail(SB200_ERR_INVALID, "at most %d spatio-temporal constraints", SB200_MAX_CONSTRAINTS);
  // SpatioTemporalConstraints::add_constraints: stable sort by epoch, dedup keeping the first
  std::vector<std::pair<int, float>> c;
  for (int i = 0; i < o.n_constraints; ++i) {
    if (!(o.constraint_max_dist[i] > 0.0f))
      return fail(SB200_ERR_INVALID, "The distance is expected to be a positive float");
    c.emplace_back(o.constraint_epochs[i], o.constraint_max_dist[i]);
  }
  std::stable_sort(c.begin(), c.end(), [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.first < b.first; });
  c.erase(std::unique(c.begin(), c.end(), [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.first == b.first; }), c.end());
  p.n_constraints = (int)c.size();
  for (size_t i = 0; i < c.size(); ++i) { p.constraint_epochs[i] = c[i].first; p.constraint_max_dist[i] = c[i].second; }
  p.is_visual = o.kind == SB200_KIND_VISUAL_SORT || o.kind == SB200_KIND_BATCH_VISUAL_SORT;
  p.is_batch = o.kind == SB200_KIND_BATCH_SORT || o.kind == SB200_KIND_BATCH_VISUAL_SORT;
  if (p.is_visual) {
    if (o.visual_kind != SB200_VIS_EUCLIDEAN && o.visual_kind != SB200_VIS_COSINE)
      return fail(SB200_ERR_INVALID, "unknown visual metric %d", o.visual_kind);
    if (o.feature_dim <= 0) return fail(SB200_ERR_INVALID, "feature_dim must be > 0 for visual trackers");
    if (o.visual_max_observations < 1 || o.visual_max_observations > sb::kMaxObs)
      return fail(SB200_ERR_CAPACITY, "visual_max_observations must be in [1, %d]", sb::kMaxObs);
    p.visual_threshold = o.visual_threshold;
    p.feature_dim = o.feature_dim;
    p.d8 = (o.feature_dim + 7) / 8 * 8;
    p.max_obs = o.visual_max_observations;
    p.min_votes = o.visual_min_votes;
    p.min_track_length = o.visual_minimal_track_length;
    p.min_area = o.visual_minimal_area;
    p.min_quality_use = o.visual_minimal_quality_use;
    p.min_quality_collect = o.visual_minimal_quality_collect;
    p.min_own_use = o.visual_minimal_own_area_percentage_use;
    p.min_own_collect = o.visual_minimal_own_area_percentage_collect;
    p.use_own_area = (o.visual_minimal_own_area_percentage_collect + o.visual_minimal_own_area_percentage_use) > 0.0f;
  } else {
    p.max_obs = 1;
    p.d8 = 8;
  }
  *out = p;
  return 0;
}

}  // namespace

struct sb200_tracker {
  sb200_options opts{};
  sb::Params P{};
  int device = 0;
  // `stream` is the tracker's own work stream.  A caller's stream (sb200_tracker_set_stream) is joined by events: the work
  // of a call is ordered after what the caller's stream held when the call was made, and that stream waits for the call's
  // frame -- the stream-order contract, without the library's kernels sitting in the caller's stream (so the library is free
  // to start the next frame's candidate preparation under the current frame, see prep_stream).
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaStream_t user_stream = nullptr;   // valid when has_user_stream (0 is the legacy default stream)
  bool has_user_stream = false;
  bool join_per_call = true;   // the caller's stream waits for every call's frame (else: sb200_stream_join)
  cudaEvent_t ev_join_req = nullptr;
  cudaEvent_t ev_user_in = nullptr, ev_user_out = nullptr;
  // candidate preparation (prep + norms / BF16 rows: needs the request only) one frame ahead, on its own stream, into one
  // of two sets of candidate-side buffers
  cudaStream_t prep_stream = nullptr;
  cudaEvent_t ev_prep_done = nullptr, ev_set_free[2]{}, ev_inputs = nullptr;
  cudaEvent_t ev_cost_done = nullptr;   // the cost kernels of the newest frame have been issued up to here (work stream)
  bool cost_done_valid = false;
  bool set_busy[2]{};
  bool prep_off = false;
  unsigned long long frame_seq = 0;
  float kernel_ms[2]{};    // screen, refine(+mode) of the last absorbed frame
  bool tc_timed = false;
  cudaStream_t copy_stream = nullptr;
  // ---- frames in flight.  predict() is stream-ordered: it enqueues a frame and returns; what only the device knows when
  // the call is made (tracks per scene, ids consumed, expired tracks) is joined in by frame_setup_kernel and read back
  // into the host mirrors when the frame is absorbed -- lazily, by a later call, or by sb200_sync().
  static constexpr int kDepth = 3;
  struct Pending {
    bool active = false;
    int n_scenes = 0, total = 0;
    std::vector<int> slots, m;
    long long live_ub = 0;     // upper bound of the records this frame's sweep appends to the wasted buffer
    HBuf h_req;                // SceneReq[n_scenes], written by the host, read by frame_setup_kernel over PCIe
    HBuf h_out;                // frame_out[n][3] | status[n] | FrameDyn, copied back at the end of the frame
    cudaEvent_t done = nullptr;
    cudaEvent_t ev[6]{}, ev_k[3]{}, ev_pos[2]{};
    bool tc_timed = false, pos_forked = false;
    int mode = 0;              // visual cost path of the frame: 0 none / exact SIMT, 1 screen + refine, 2 dense tensor-core
  } pend[kDepth];
  int pend_head = 0, pend_count = 0;
  std::vector<int> pending_add;   // per slot: detections of the frames in flight (each can add at most that many tracks)
  long long inflight_live_ub = 0;
  int async_rc = 0;
  std::string async_err;
  // cumulative work counters over the absorbed frames (bench.py reads them around its timed region)
  unsigned long long acc_units_mn = 0, acc_units_rows = 0, acc_frames = 0;
  double host_ms_total = 0.0, host_ms_blocked = 0.0;   // wall time inside predict(), and the part of it spent waiting for the device
  unsigned long long host_calls = 0;
  double acc_stage_ms[5]{}, acc_kernel_ms[2]{};
  unsigned long long acc_tc_frames = 0;
  // side stream of the positional stage (visual trackers): the culled scan runs next to the refinement of the visual
  // survivors instead of in front of the screen
  cudaStream_t pos_stream = nullptr;   // side stream: frame tables beside the candidate preparation, sweep beside the feature store
  bool side_off = false;
  cudaEvent_t ev_fork[2]{}, ev_join = nullptr;
  float stage_ms[5]{};
  // scene table
  std::unordered_map<uint64_t, int> slot_of;
  std::vector<uint64_t> scene_of_slot;
  std::vector<uint32_t> epoch;      // host epoch db
  std::vector<int> n_tracks;        // host mirror: live tracks per slot (expired tracks leave the device store every frame)
  std::vector<int> n_hidden;        // expired tracks per slot swept early and not yet collected in the reference's sense
  std::vector<int> arena_top;       // host mirror: feature blocks handed out per slot (rows the screen scans = top * K)
  std::vector<uint64_t> last_req_scenes;   // scene ids of the previous request and their slots (steady-state fast path)
  std::vector<int> last_req_slots;
  int64_t revealed = 0;             // wasted-buffer records [0, revealed) are collected; [revealed, wasted_count) hidden
  int scene_cap = 0, track_cap = 0;
  DBuf b_idc;                       // device id counter (ids consumed so far)
  int auto_waste_counter = 100, auto_waste_periodicity = 100;
  int64_t wasted_count = 0;
  // device track store
  sb::TrackStore ts{};
  DBuf b_id, b_epoch, b_length, b_custom, b_vt, b_pred, b_obs, b_radius, b_kst, b_vert, b_hpred, b_hobs, b_feat, b_feat_bf16, b_fnorm2, b_obs_phys,
      b_obs_hasf, b_obs_q, b_obs_n, b_feat_cnt, b_fblk, b_blk_owner, b_blk_free;
  DBuf b_ntracks, b_cur_epoch, b_scene_ids, b_nfree, b_atop;
  // wasted
  sb::WastedBuf wb{};
  DBuf w_count, w_id, w_scene, w_epoch, w_length, w_pred, w_obs, w_hpred, w_hobs;
  int hist_len = 1;   // boxes of history kept per track (1: only the last ones, the SortTrack columns)
  // frame buffers
  // two input staging sets: sb200_prefetch_inputs() fills one while the kernels of the previous frame read the other
  struct Staging {
    DBuf boxes, feat, hasf, quality, custom, own;
    const void* key[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // the six host columns of the prefetch
    int total = -1;
    bool pending = false;   // holds a prefetched request that no predict call has consumed yet
    cudaEvent_t ev = nullptr;
    cudaEvent_t ev0 = nullptr;   // start of the set's prefetch copy (SB200_TRACE timing)
    cudaEvent_t ev_read = nullptr;   // end of the last frame that read this set (a prefetch into it waits for that)
    bool read_pending = false;
  } stg[2];
  int stg_last = 1;   // staging set used by the most recent predict
  DBuf f_cbox2[2], f_cradius2[2], f_cconf2[2], f_cvert2[2], f_cflags2[2], f_cnorm22[2], f_cbf162[2], f_decided2[2];   // candidate side, two sets
  DBuf f_winner, f_cvt, f_pos, f_vis, f_scenes, f_newcount,
      f_status, f_featdst, f_apprank, f_appmeta, f_posgq, f_frameout, f_excl, f_prewin, f_own, f_ownovf, f_dyn, f_ws, f_tmeta, f_rowinfo, f_slabc, f_slabm, f_slabmask,
      f_dscene, f_maxc, f_maxcval, f_drowb, f_dcolb, f_slabk, f_scene_max, f_tiles, f_pairs, f_colmeta, f_colgeo, f_colb, f_colvalid, f_rowmeta, f_poslist, f_counters, f_visval;
  int num_sms = 148;
  DBuf o_ids, o_epochs, o_lengths, o_vt, o_pred, o_obs;
  HBuf h_small;
  bool adapt_dense = false;     // a nominally selective threshold whose survivor lists overflow: treat it as non-selective
  unsigned long long acc_dense_scenes = 0;   // scenes the exact SIMT fallback had to take (over all absorbed frames)
  int last_dense_scenes = 0;
  bool seen_features = false;   // a request has carried feature rows (the feature dimension is fixed from then on)
  int last_n_scenes = 0;   // scenes of the last frame (sb200_last_costs reads its scene table back from the device)

  ~sb200_tracker() {
    cudaSetDevice(device);
    DBuf* all[] = {&b_id, &b_epoch, &b_length, &b_custom, &b_vt, &b_pred, &b_obs, &b_radius, &b_kst, &b_vert, &b_hpred, &b_hobs, &w_hpred, &w_hobs, &b_feat,
                   &b_feat_bf16, &f_scene_max, &f_tiles, &f_pairs, &f_colmeta, &f_colgeo, &f_colb, &f_colvalid, &f_rowmeta, &f_poslist, &f_counters, &f_visval, &b_fnorm2, &b_obs_phys, &b_obs_hasf, &b_obs_q, &b_obs_n, &b_feat_cnt, &b_ntracks, &b_cur_epoch,
                   &b_fblk, &b_blk_owner, &b_blk_free, &b_nfree, &b_atop, &f_frameout, &f_excl, &f_prewin, &f_own, &f_ownovf, &f_dyn, &b_idc, &f_ws, &f_tmeta, &f_rowinfo, &f_slabc, &f_slabm, &f_slabmask, &f_dscene, &f_maxc, &f_maxcval, &f_drowb, &f_dcolb, &f_slabk,
                   &b_scene_ids, &w_count, &w_id, &w_scene, &w_epoch, &w_length, &w_pred, &w_obs, &f_winner, &f_cvt, &f_pos, &f_vis, &f_scenes, &f_newcount, &f_status,
                   &f_featdst, &f_apprank, &f_appmeta, &f_posgq, &o_ids, &o_epochs, &o_lengths, &o_vt, &o_pred, &o_obs};
    for (DBuf* b : all) b->release();
    for (int k = 0; k < 2; ++k) {
      f_cbox2[k].release(); f_cradius2[k].release(); f_cconf2[k].release(); f_cvert2[k].release(); f_cflags2[k].release();
      f_cnorm22[k].release(); f_cbf162[k].release(); f_decided2[k].release();
    }
    if (stream) cudaStreamSynchronize(stream);
    h_small.release();
    for (auto& g : stg) {
      g.boxes.release(); g.feat.release(); g.hasf.release(); g.quality.release(); g.custom.release(); g.own.release();
      if (g.ev) cudaEventDestroy(g.ev);
      if (g.ev0) cudaEventDestroy(g.ev0);
      if (g.ev_read) cudaEventDestroy(g.ev_read);
    }
    for (auto& q : pend) {
      q.h_req.release(); q.h_out.release();
      if (q.done) cudaEventDestroy(q.done);
      for (auto& e : q.ev) if (e) cudaEventDestroy(e);
      for (auto& e : q.ev_k) if (e) cudaEventDestroy(e);
      for (auto& e : q.ev_pos) if (e) cudaEventDestroy(e);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (pos_stream) cudaStreamDestroy(pos_stream);
    if (prep_stream) cudaStreamDestroy(prep_stream);
    for (cudaEvent_t e : {ev_user_in, ev_user_out, ev_prep_done, ev_set_free[0], ev_set_free[1], ev_inputs, ev_join_req, ev_cost_done}) if (e) cudaEventDestroy(e);
    for (auto& e : ev_fork) if (e) cudaEventDestroy(e);
    if (ev_join) cudaEventDestroy(ev_join);
    if (own_stream && stream) cudaStreamDestroy(stream);
  }

  // (re)allocates the track store for scene_cap x track_cap rows, preserving the live rows
  template <typename T>
  int regrow(DBuf& b, T** field, int width, int new_scenes, int new_tracks, bool zero = false) {
    size_t row = (size_t)width * sizeof(T);
    DBuf nb;
    int rc = nb.ensure(std::max<size_t>(1, (size_t)new_scenes * new_tracks * row));
    if (rc) return rc;
    // rows that are moved whole although only partly written (the history rings of short tracks) start out defined
    if (zero && cudaMemsetAsync(nb.p, 0, std::max<size_t>(1, (size_t)new_scenes * new_tracks * row), stream) != cudaSuccess)
      return fail(SB200_ERR_CUDA, "store regrow memset failed");
    if (b.p && scene_cap > 0 && track_cap > 0) {
      cudaError_t e = cudaMemcpy2DAsync(nb.p, (size_t)new_tracks * row, b.p, (size_t)track_cap * row,
                                        (size_t)std::min(track_cap, new_tracks) * row, (size_t)scene_cap,
                                        cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) return fail(SB200_ERR_CUDA, "store regrow copy failed: %s", cudaGetErrorString(e));
      e = cudaStreamSynchronize(stream);
      if (e != cudaSuccess) return fail(SB200_ERR_CUDA, "store regrow sync failed: %s", cudaGetErrorString(e));
    }
    b.release();
    b = nb;
    *field = b.as<T>();
    return 0;
  }

  int ensure_store(int need_scenes, int need_tracks) {
    if (need_scenes <= scene_cap && need_tracks <= track_cap) return 0;
    int ns = scene_cap, nt = track_cap;
    if (need_scenes > ns) ns = std::max(need_scenes, std::max(4, ns * 2));
    if (need_tracks > nt) nt = std::max(need_tracks, std::max(64, nt * 2));
    const int K = P.max_obs;
    int rc = 0;
    if ((rc = regrow(b_id, &ts.id, 1, ns, nt))) return rc;
    if ((rc = regrow(b_epoch, &ts.epoch, 1, ns, nt))) return rc;
    if ((rc = regrow(b_length, &ts.length, 1, ns, nt))) return rc;
    if ((rc = regrow(b_custom, &ts.custom, 1, ns, nt))) return rc;
    if ((rc = regrow(b_vt, &ts.vt, 1, ns, nt))) return rc;
    if ((rc = regrow(b_pred, &ts.pred, 6, ns, nt))) return rc;
    if ((rc = regrow(b_obs, &ts.obs, 6, ns, nt))) return rc;
    if ((rc = regrow(b_radius, &ts.radius, 1, ns, nt))) return rc;
    if ((rc = regrow(b_kst, &ts.kst, sb::kStateStride, ns, nt))) return rc;
    ts.kst_stride = sb::kStateStride;
    if (P.positional_kind == SB200_POS_IOU)
      if ((rc = regrow(b_vert, &ts.vert, 8, ns, nt))) return rc;
    if (hist_len > 1) {
      if ((rc = regrow(b_hpred, &ts.hist_pred, 6 * hist_len, ns, nt, true))) return rc;
      if ((rc = regrow(b_hobs, &ts.hist_obs, 6 * hist_len, ns, nt, true))) return rc;
      ts.hist_len = hist_len;
    }
    if (P.is_visual) {
      if ((rc = regrow(b_feat, &ts.feat, K * P.d8, ns, nt))) return rc;
      {
        unsigned short* tmp = reinterpret_cast<unsigned short*>(ts.feat_bf16);
        if ((rc = regrow(b_feat_bf16, &tmp, K * P.d8, ns, nt))) return rc;
        ts.feat_bf16 = tmp;
      }
      if ((rc = regrow(b_fnorm2, &ts.fnorm2, K, ns, nt))) return rc;
      if ((rc = regrow(b_obs_phys, &ts.obs_phys, K, ns, nt))) return rc;
      if ((rc = regrow(b_obs_hasf, &ts.obs_hasf, K, ns, nt))) return rc;
      if ((rc = regrow(b_obs_q, &ts.obs_q, K, ns, nt))) return rc;
      if ((rc = regrow(b_obs_n, &ts.obs_n, 1, ns, nt))) return rc;
      if ((rc = regrow(b_feat_cnt, &ts.feat_cnt, 1, ns, nt))) return rc;
      if ((rc = regrow(b_fblk, &ts.fblk, 1, ns, nt))) return rc;
      if ((rc = regrow(b_blk_owner, &ts.blk_owner, 1, ns, nt))) return rc;
      if ((rc = regrow(b_blk_free, &ts.blk_free, 1, ns, nt))) return rc;
    }
    if (ns != scene_cap) {
      // per-slot small arrays
      DBuf n1, n2, n3, n4, n5;
      if ((rc = n1.ensure(sizeof(int) * ns))) return rc;
      if ((rc = n2.ensure(sizeof(unsigned int) * ns))) return rc;
      if ((rc = n3.ensure(sizeof(unsigned long long) * ns))) return rc;
      if ((rc = n4.ensure(sizeof(int) * ns))) return rc;
      if ((rc = n5.ensure(sizeof(int) * ns))) return rc;
      CU(cudaMemsetAsync(n1.p, 0, sizeof(int) * ns, stream));
      CU(cudaMemsetAsync(n4.p, 0, sizeof(int) * ns, stream));
      CU(cudaMemsetAsync(n5.p, 0, sizeof(int) * ns, stream));
      if (b_ntracks.p && scene_cap > 0) {
        CU(cudaMemcpyAsync(n1.p, b_ntracks.p, sizeof(int) * scene_cap, cudaMemcpyDeviceToDevice, stream));
        CU(cudaMemcpyAsync(n4.p, b_nfree.p, sizeof(int) * scene_cap, cudaMemcpyDeviceToDevice, stream));
        CU(cudaMemcpyAsync(n5.p, b_atop.p, sizeof(int) * scene_cap, cudaMemcpyDeviceToDevice, stream));
      }
      CU(cudaStreamSynchronize(stream));
      b_ntracks.release(); b_cur_epoch.release(); b_scene_ids.release(); b_nfree.release(); b_atop.release();
      b_ntracks = n1; b_cur_epoch = n2; b_scene_ids = n3; b_nfree = n4; b_atop = n5;
      ts.n_free = b_nfree.as<int>();
      ts.arena_top = b_atop.as<int>();
    }
    scene_cap = ns;
    track_cap = nt;
    ts.track_cap = nt;
    return 0;
  }

  int slot_for(uint64_t scene_id, bool create) {
    auto it = slot_of.find(scene_id);
    if (it != slot_of.end()) return it->second;
    if (!create) return -1;
    int s = (int)scene_of_slot.size();
    slot_of[scene_id] = s;
    scene_of_slot.push_back(scene_id);
    epoch.push_back(0);
    n_tracks.push_back(0);
    n_hidden.push_back(0);
    arena_top.push_back(0);
    pending_add.push_back(0);
    return s;
  }

  int ensure_wasted(int64_t need) {
    if (need <= wb.cap) return 0;
    int64_t ncap = std::max<int64_t>(need, std::max<int64_t>(1024, (int64_t)wb.cap * 3));
    // wasted records are drained by sb200_wasted; growing preserves the pending ones
    DBuf nid, nsc, nep, nle, npr, nob, nhp, nho;
    int rc;
    if ((rc = nid.ensure(8 * ncap)) || (rc = nsc.ensure(8 * ncap)) || (rc = nep.ensure(4 * ncap)) ||
        (rc = nle.ensure(4 * ncap)) || (rc = npr.ensure(24 * ncap)) || (rc = nob.ensure(24 * ncap)))
      return rc;
    const size_t hrow = (size_t)24 * hist_len;
    if (hist_len > 1 && ((rc = nhp.ensure(hrow * ncap)) || (rc = nho.ensure(hrow * ncap)))) return rc;
    if (wasted_count > 0) {
      CU(cudaMemcpyAsync(nid.p, w_id.p, 8 * wasted_count, cudaMemcpyDeviceToDevice, stream));
      CU(cudaMemcpyAsync(nsc.p, w_scene.p, 8 * wasted_count, cudaMemcpyDeviceToDevice, stream));
      CU(cudaMemcpyAsync(nep.p, w_epoch.p, 4 * wasted_count, cudaMemcpyDeviceToDevice, stream));
      CU(cudaMemcpyAsync(nle.p, w_length.p, 4 * wasted_count, cudaMemcpyDeviceToDevice, stream));
      CU(cudaMemcpyAsync(npr.p, w_pred.p, 24 * wasted_count, cudaMemcpyDeviceToDevice, stream));
      CU(cudaMemcpyAsync(nob.p, w_obs.p, 24 * wasted_count, cudaMemcpyDeviceToDevice, stream));
      if (hist_len > 1) {
        CU(cudaMemcpyAsync(nhp.p, w_hpred.p, hrow * wasted_count, cudaMemcpyDeviceToDevice, stream));
        CU(cudaMemcpyAsync(nho.p, w_hobs.p, hrow * wasted_count, cudaMemcpyDeviceToDevice, stream));
      }
      CU(cudaStreamSynchronize(stream));
    }
    w_id.release(); w_scene.release(); w_epoch.release(); w_length.release(); w_pred.release(); w_obs.release();
    w_hpred.release(); w_hobs.release();
    w_id = nid; w_scene = nsc; w_epoch = nep; w_length = nle; w_pred = npr; w_obs = nob; w_hpred = nhp; w_hobs = nho;
    wb.hist_pred = w_hpred.as<float>();
    wb.hist_obs = w_hobs.as<float>();
    if (!w_count.p) {
      if ((rc = w_count.ensure(sizeof(int)))) return rc;
      CU(cudaMemsetAsync(w_count.p, 0, sizeof(int), stream));
    }
    wb.cap = (int)ncap;
    wb.count = w_count.as<int>();
    wb.id = w_id.as<unsigned long long>();
    wb.scene = w_scene.as<unsigned long long>();
    wb.epoch = w_epoch.as<unsigned int>();
    wb.length = w_length.as<unsigned int>();
    wb.pred = w_pred.as<float>();
    wb.obs = w_obs.as<float>();
    return 0;
  }

  // TrackerAPI::auto_waste (src/trackers/tracker_api.rs:81-88)
  int run_waste() {
    int n_slots = (int)scene_of_slot.size();
    if (n_slots == 0) return 0;
    int64_t active = 0;
    int max_n = 0;
    for (int v : n_tracks) { active += v; max_n = std::max(max_n, v); }
    if (active == 0) {
      revealed = wasted_count;
      std::fill(n_hidden.begin(), n_hidden.end(), 0);
      return 0;
    }
    int rc = ensure_wasted(wasted_count + active);
    if (rc) return rc;
    if ((rc = h_small.ensure((size_t)n_slots * 16))) return rc;
    unsigned int* he = h_small.as<unsigned int>();
    for (int s = 0; s < n_slots; ++s) he[s] = epoch[s];
    CU(cudaMemcpyAsync(b_cur_epoch.p, he, sizeof(unsigned int) * n_slots, cudaMemcpyHostToDevice, stream));
    CU(cudaStreamSynchronize(stream));
    unsigned long long* hs = h_small.as<unsigned long long>();
    for (int s = 0; s < n_slots; ++s) hs[s] = scene_of_slot[s];
    CU(cudaMemcpyAsync(b_scene_ids.p, hs, sizeof(unsigned long long) * n_slots, cudaMemcpyHostToDevice, stream));
    sb::launch_waste(P, ts, n_slots, b_cur_epoch.as<unsigned int>(), b_scene_ids.as<unsigned long long>(),
                     b_ntracks.as<int>(), wb, max_n, stream);
    CU(cudaGetLastError());
    int* hn = h_small.as<int>();
    CU(cudaStreamSynchronize(stream));
    CU(cudaMemcpyAsync(hn, b_ntracks.p, sizeof(int) * n_slots, cudaMemcpyDeviceToHost, stream));
    int hc = 0;
    CU(cudaMemcpyAsync(&hc, w_count.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    for (int s = 0; s < n_slots; ++s) n_tracks[s] = hn[s];
    wasted_count = hc;
    // this is one of the reference's collection points: everything swept early becomes visible now
    revealed = wasted_count;
    std::fill(n_hidden.begin(), n_hidden.end(), 0);
    return 0;
  }

  // removes the first n records of the wasted buffer (the rest, hidden ones included, shift to the front)
  int drop_wasted_front(int64_t n) {
    if (n <= 0 || !w_count.p) return 0;
    n = std::min(n, wasted_count);
    const int64_t rest = wasted_count - n;
    cudaStream_t st = stream;
    int rc = 0;
    if (rest > 0) {
      DBuf tmp;   // overlapping device-to-device moves are done through a temporary
      if ((rc = tmp.ensure((size_t)rest * 24))) return rc;
      auto shift = [&](void* base, size_t el) -> int {
        CU(cudaMemcpyAsync(tmp.p, (char*)base + n * el, rest * el, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(base, tmp.p, rest * el, cudaMemcpyDeviceToDevice, st));
        return 0;
      };
      if ((rc = shift(wb.id, 8)) || (rc = shift(wb.scene, 8)) || (rc = shift(wb.epoch, 4)) || (rc = shift(wb.length, 4)) ||
          (rc = shift(wb.pred, 24)) || (rc = shift(wb.obs, 24)))
        return rc;
      if (hist_len > 1) {
        tmp.release();
        if ((rc = tmp.ensure((size_t)rest * 24 * hist_len))) return rc;
        if ((rc = shift(wb.hist_pred, (size_t)24 * hist_len)) || (rc = shift(wb.hist_obs, (size_t)24 * hist_len))) return rc;
      }
      CU(cudaStreamSynchronize(st));
      tmp.release();
    }
    int newc = (int)rest;
    CU(cudaMemcpyAsync(w_count.p, &newc, sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    wasted_count = rest;
    revealed = std::max<int64_t>(0, revealed - n);
    return 0;
  }

  int predict(int32_t n_scenes, const uint64_t* scene_ids, const int32_t* det_offsets, const float* boxes,
              const float* features, const uint8_t* has_feature, const float* quality, const int64_t* custom_ids,
              const float* own_area, const sb200_predict_out* out, bool device_io, bool wait);
  int absorb_oldest(bool block);
  int poll();
  int drain(const char* why = nullptr);
  int ens(DBuf& b, size_t need, const char* name = "") {   // ensure() that never reallocates under a frame in flight
    if (need <= b.bytes) return 0;
    static const bool trace_ens = getenv("SB200_TRACE") != nullptr;
    if (trace_ens) fprintf(stderr, "[sb200] %s grows: %zu -> %zu bytes\n", name, b.bytes, need);
    int rc = drain("frame buffer grows");
    if (rc) return rc;
    return b.ensure(need);
  }
  static size_t dyn_offset(int n_scenes) { return ((size_t)n_scenes * 16 + 15) / 16 * 16; }
  static constexpr int kDenseVoteCap = 8192;   // visual entries per scene of the voting kernel on the dense path (power of two)
};

// Reads back what the oldest frame in flight left for the host: per scene {live tracks, arena blocks, newly expired} and
// the status word, plus the frame scalars.  block == false: only if the frame has completed.  Returns 1 when nothing was
// absorbed.  Errors of an asynchronous frame are kept (async_rc / async_err) for the next call that reports status.
int sb200_tracker::absorb_oldest(bool block) {
  if (pend_count == 0) return 1;
  Pending& q = pend[pend_head];
  cudaError_t e;
  if (block) {
    const auto w0 = std::chrono::steady_clock::now();
    e = cudaEventSynchronize(q.done);
    host_ms_blocked += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
  } else {
    e = cudaEventQuery(q.done);
  }
  if (e == cudaErrorNotReady) { cudaGetLastError(); return 1; }
  const int n = q.n_scenes;
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (!async_rc) { async_rc = SB200_ERR_CUDA; async_err = std::string("a frame in flight failed: ") + cudaGetErrorString(e); }
  } else {
    const int* h_fo = q.h_out.as<int>();            // [n][3] live tracks, arena blocks, newly expired
    const int* h_status = h_fo + 3 * (size_t)n;
    const sb::FrameDyn* dyn = reinterpret_cast<const sb::FrameDyn*>(reinterpret_cast<const char*>(q.h_out.p) + dyn_offset(n));
    for (int s = 0; s < n; ++s) {
      const int slot = q.slots[s];
      if (h_status[s] && !async_rc) {
        async_rc = (h_status[s] & 2) ? SB200_ERR_CAPACITY : SB200_ERR_INTERNAL;
        char buf[160];
        snprintf(buf, sizeof(buf), (h_status[s] & 2) ? "more than 2800 boxes overlap one detection of scene %llu (own-area shares)"
                                                      : "track store overflow in scene %llu",
                 (unsigned long long)scene_of_slot[slot]);
        async_err = buf;
      }
      n_tracks[slot] = h_fo[3 * s];
      arena_top[slot] = h_fo[3 * s + 1];
      n_hidden[slot] += h_fo[3 * s + 2];   // swept from the device store, not yet collected in the reference's sense
      wasted_count += h_fo[3 * s + 2];
    }
    acc_units_mn += dyn->units_mn;
    acc_units_rows += dyn->units_rows;
    acc_frames += 1;
    {
      const int dense_scenes = *reinterpret_cast<const int*>(reinterpret_cast<const char*>(dyn) + sizeof(sb::FrameDyn));
      if (q.mode != 0) acc_dense_scenes += (unsigned long long)dense_scenes;   // mode 0 IS the exact kernel: not a fallback
      last_dense_scenes = dense_scenes;
      // most scenes of a screened frame overflowed their survivor lists: the threshold cuts (almost) nothing, so the
      // following frames take the dense tensor-core path; and back, if that path's precondition keeps failing
      if (q.mode == 1 && n >= 1 && dense_scenes * 4 > n) adapt_dense = true;
      if (q.mode == 2 && n >= 1 && dense_scenes * 4 > n) adapt_dense = false;
    }
    for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&stage_ms[i], q.ev[i], q.ev[i + 1]);
    if (q.pos_forked) {
      // lazy frame: the culled scan sits inside the visual span (after the BestFit pre-pass): report it on its own
      float fill_ms = stage_ms[1], scan_ms = 0.0f;
      cudaEventElapsedTime(&scan_ms, q.ev_pos[0], q.ev_pos[1]);
      stage_ms[1] = scan_ms;
      stage_ms[2] += fill_ms - scan_ms;
    }
    tc_timed = q.tc_timed;
    kernel_ms[0] = kernel_ms[1] = 0.0f;
    if (tc_timed) { cudaEventElapsedTime(&kernel_ms[0], q.ev_k[0], q.ev_k[1]); cudaEventElapsedTime(&kernel_ms[1], q.ev_k[1], q.ev_k[2]); }
    for (int i = 0; i < 5; ++i) acc_stage_ms[i] += stage_ms[i];
    if (tc_timed) { acc_kernel_ms[0] += kernel_ms[0]; acc_kernel_ms[1] += kernel_ms[1]; acc_tc_frames += 1; }
    static const bool trace_abs = getenv("SB200_TRACE") != nullptr;
    if (trace_abs)
      fprintf(stderr, "[sb200] frame absorbed: mode %d, %d of %d scenes on the exact SIMT fallback; prep %.3f pos %.3f vis %.3f vote %.3f apply %.3f ms (main kernel %.3f, refine %.3f)\n",
              q.mode, last_dense_scenes, n, stage_ms[0], stage_ms[1], stage_ms[2], stage_ms[3], stage_ms[4], kernel_ms[0], kernel_ms[1]);
    cudaGetLastError();   // an event that was never recorded in this frame leaves cudaErrorInvalidResourceHandle behind
  }
  for (int s = 0; s < n; ++s) pending_add[q.slots[s]] -= q.m[s];
  inflight_live_ub -= q.live_ub;
  q.active = false;
  pend_head = (pend_head + 1) % kDepth;
  pend_count -= 1;
  return 0;
}

int sb200_tracker::poll() {
  while (pend_count > 0 && absorb_oldest(false) == 0) {}
  return 0;
}

int sb200_tracker::drain(const char* why) {
  static const bool trace_dr = getenv("SB200_TRACE") != nullptr;
  if (trace_dr && why && pend_count > 0) fprintf(stderr, "[sb200] predict waits for %d frame(s) in flight: %s\n", pend_count, why);
  while (pend_count > 0) absorb_oldest(true);
  if (async_rc) {
    const int rc = async_rc;
    async_rc = 0;
    return fail(rc, "%s", async_err.c_str());
  }
  return 0;
}

#define ENS(b, ...) ens(b, __VA_ARGS__, #b)
int sb200_tracker::predict(int32_t n_scenes, const uint64_t* scene_ids, const int32_t* det_offsets, const float* boxes,
                           const float* features, const uint8_t* has_feature, const float* quality,
                           const int64_t* custom_ids, const float* own_area, const sb200_predict_out* out,
                           bool device_io, bool wait) {
  CU(cudaSetDevice(device));
  static const bool trace = getenv("SB200_TRACE") != nullptr;
  auto tnow = [] { return std::chrono::steady_clock::now(); };
  auto t_begin = tnow();
  auto since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(tnow() - a).count(); };
  if (n_scenes < 0) return fail(SB200_ERR_INVALID, "n_scenes must be >= 0");
  if (n_scenes > 0 && (!scene_ids || !det_offsets)) return fail(SB200_ERR_INVALID, "scene_ids / det_offsets are NULL");
  const int total = n_scenes > 0 ? det_offsets[n_scenes] : 0;
  if (n_scenes > 0 && det_offsets[0] != 0) return fail(SB200_ERR_INVALID, "det_offsets[0] must be 0");
  for (int s = 0; s < n_scenes; ++s)
    if (det_offsets[s + 1] < det_offsets[s]) return fail(SB200_ERR_INVALID, "det_offsets must be non-decreasing");
  if (total > 0 && !boxes) return fail(SB200_ERR_INVALID, "boxes is NULL");
  if (!P.is_visual) { features = nullptr; has_feature = nullptr; quality = nullptr; own_area = nullptr; }
  int rc = 0;
  // frames that have completed since the last call hand over their results; an error of an earlier asynchronous frame
  // is reported now
  poll();
  if (async_rc) { if ((rc = drain("error of an earlier frame"))) return rc; }
  // same scene list as the previous request (the steady state of a batch tracker): validated slots are reused
  const bool same_req = (int)last_req_scenes.size() == n_scenes && n_scenes > 0 &&
                        memcmp(last_req_scenes.data(), scene_ids, sizeof(uint64_t) * (size_t)n_scenes) == 0;
  if (!same_req && n_scenes > 0) {
    std::unordered_map<uint64_t, int> seen;
    for (int s = 0; s < n_scenes; ++s)
      if (!seen.emplace(scene_ids[s], s).second) return fail(SB200_ERR_INVALID, "scene %llu appears twice in one request", (unsigned long long)scene_ids[s]);
  }
  // Everything that can refuse the request is checked BEFORE any tracker state changes (epochs, the auto-waste counter):
  // the on-chip assignment solver holds a scene's rows and columns in shared memory.
  {
    int max_m0 = 0, max_n0 = 0;
    for (int s = 0; s < n_scenes; ++s) {
      const int m = det_offsets[s + 1] - det_offsets[s];
      int slot = -1;
      if (same_req) slot = last_req_slots[s];
      else { auto it = slot_of.find(scene_ids[s]); if (it != slot_of.end()) slot = it->second; }
      const int nub = slot >= 0 ? n_tracks[slot] + pending_add[slot] : 0;
      max_m0 = std::max(max_m0, m);
      max_n0 = std::max(max_n0, nub);
    }
    const int cap0 = P.is_visual ? kDenseVoteCap : 0;
    if (sb::voting_smem_need(max_m0, max_n0, cap0) > sb::kVotingSmemLimit) {
      if (pend_count > 0) {   // the bound counts every detection in flight as a new track: get the exact counts first
        if ((rc = drain("assignment solver bound"))) return rc;
        max_n0 = 0;
        for (int s = 0; s < n_scenes; ++s) {
          auto it = slot_of.find(scene_ids[s]);
          if (it != slot_of.end()) max_n0 = std::max(max_n0, n_tracks[it->second]);
        }
      }
      if (sb::voting_smem_need(max_m0, max_n0, cap0) > sb::kVotingSmemLimit)
        return fail(SB200_ERR_CAPACITY, "scene too large for the on-chip assignment solver (m=%d, n=%d)", max_m0, max_n0);
    }
  }
  // auto-waste tick (src/trackers/sort/simple_api.rs:115-120): a collection point of the reference, host and device meet
  if (auto_waste_counter == 0) {
    if ((rc = drain("auto-waste tick"))) return rc;
    if ((rc = run_waste())) return rc;
    auto_waste_counter = auto_waste_periodicity;
  } else auto_waste_counter -= 1;
  if (n_scenes == 0) return wait ? drain() : 0;

  if (!same_req) {
    last_req_slots.resize(n_scenes);
    for (int s = 0; s < n_scenes; ++s) last_req_slots[s] = slot_for(scene_ids[s], true);
    last_req_scenes.assign(scene_ids, scene_ids + n_scenes);
  }
  if (pend_count == kDepth) {
    if (trace) fprintf(stderr, "[sb200] predict waits: ring of %d frames full\n", kDepth);
    absorb_oldest(true);
  }   // back-pressure: at most kDepth frames in flight

  // ---- upper bounds of everything the device will size exactly (tracks per scene with frames still in flight)
  const int K = P.max_obs;
  std::vector<int> n_ub(n_scenes), nb_ub(n_scenes), m_of(n_scenes);
  int max_m = 0, max_n = 0, max_nb = 0, need_tracks = 0;
  long long live_ub = 0, pos_total = 0, vis_total = 0, col_total = 0, posl_total = 0, visl_total = 0, work = 0;
  auto bounds = [&]() {
    max_m = max_n = max_nb = need_tracks = 0;
    live_ub = pos_total = vis_total = col_total = work = 0;
    for (int s = 0; s < n_scenes; ++s) {
      const int slot = last_req_slots[s];
      const int m = det_offsets[s + 1] - det_offsets[s];
      m_of[s] = m;
      n_ub[s] = n_tracks[slot] + pending_add[slot];
      nb_ub[s] = P.is_visual ? arena_top[slot] + pending_add[slot] : 0;
      live_ub += n_ub[s];
      pos_total += (long long)m * n_ub[s];
      if (P.is_visual) vis_total += (long long)m * n_ub[s] * K;
      col_total += ((long long)nb_ub[s] * K + 127) / 128 * 128;
      work += (long long)m * n_ub[s] * K;
      max_m = std::max(max_m, m);
      max_n = std::max(max_n, n_ub[s]);
      max_nb = std::max(max_nb, nb_ub[s]);
      need_tracks = std::max(need_tracks, n_ub[s] + m);
    }
  };
  bounds();
  {
    // Capacity has to hold the UPPER BOUNDS: with frames in flight every queued detection counts as a possible new track.
    // So the store is sized for the pipeline once -- the caller's hint (or what this frame needs) plus the detections of
    // kDepth frames -- and later frames neither wait for the device nor reallocate.
    const int pipe_room = kDepth * std::max(max_m, opts.max_dets_per_scene_hint);
    int hint_s = std::max((int)scene_of_slot.size(), opts.max_scenes_hint);
    if (hint_s > scene_cap || need_tracks > track_cap) {
      // the store has to grow: meet the device first (the exact counts are the ones to grow from)
      if ((rc = drain("track store bound"))) return rc;
      bounds();
      // grow generously (a regrow copies the whole feature arena: tens of milliseconds): what is needed now plus the
      // pipeline's room, and at least half again as much as before
      const int want_t = std::max(need_tracks, opts.max_tracks_per_scene_hint) + pipe_room;
      if (hint_s > scene_cap || need_tracks > track_cap)
        if ((rc = ensure_store(hint_s, std::max(want_t, track_cap + track_cap / 2)))) return rc;
    }
    // room for every live track of the frames in flight and of this one in the wasted buffer: the end-of-frame sweep
    // appends without a host check
    if (wasted_count + inflight_live_ub + live_ub + 1 > wb.cap) {
      // the bound that failed is the pipeline's (frames in flight + this one): grow for twice that, or the next frame meets
      // the same bound again -- after the wait the exact counts alone would fit and nothing would grow
      const long long pipe_need = inflight_live_ub + live_ub;
      if ((rc = drain("wasted buffer bound"))) return rc;
      bounds();
      // a full ring: kDepth frames in flight plus this one, each bounded by the live tracks plus every detection queued before it
      const long long dets = std::max<long long>(total, (long long)std::max(opts.max_scenes_hint, n_scenes) * opts.max_dets_per_scene_hint);
      const long long ring_need = (long long)(kDepth + 1) * (live_ub + (long long)kDepth * dets);
      // and the records already waiting for collection doubled: a caller that collects rarely pays for few regrows
      // and several times that while it is cheap (<= ~2 GB of records): a regrow drains the ring and reallocates
      const long long rec_bytes = 72 + (hist_len > 1 ? 48ll * hist_len : 0);
      const long long base_need = std::max<long long>(ring_need, 2 * pipe_need);
      const long long mult = std::max<long long>(1, std::min<long long>(8, (2ll << 30) / std::max<long long>(1, base_need * rec_bytes)));
      if ((rc = ensure_wasted(2 * wasted_count + mult * base_need + 1))) return rc;
    }
    if (!b_idc.p) {
      if ((rc = b_idc.ensure(8))) return rc;
      CU(cudaMemsetAsync(b_idc.p, 0, 8, stream));
    }
  }
  // ---- this frame's slot in the ring
  Pending& q = pend[(pend_head + pend_count) % kDepth];
  if (!q.done) {
    CU(cudaEventCreateWithFlags(&q.done, cudaEventDisableTiming));
    for (auto& e : q.ev) CU(cudaEventCreate(&e));
    for (auto& e : q.ev_k) CU(cudaEventCreate(&e));
    for (auto& e : q.ev_pos) CU(cudaEventCreate(&e));
  }
  if ((rc = q.h_req.ensure(sizeof(sb::SceneReq) * (size_t)n_scenes)) ||
      (rc = q.h_out.ensure(dyn_offset(n_scenes) + sizeof(sb::FrameDyn) + 16)))
    return rc;
  // visual cost path of this frame: tensor-core screen + exact refinement for large frames with a selective threshold, the
  // dense tensor-core weight sums for thresholds that cut nothing, the exact SIMT kernel otherwise (small frames, and as the
  // device-side fallback of single scenes)
  bool want_tc = false, want_dense = false;
  if (P.is_visual && features != nullptr && total > 0) {
    const bool selective = P.visual_kind == SB200_VIS_EUCLIDEAN ? (P.visual_threshold < 1e18f) : (P.visual_threshold > -1.0f);
    const bool big = P.d8 >= 64 && work * P.d8 >= (1ll << 28);
    const bool dense_ok = P.n_constraints == 0;
    want_tc = selective && big;
    want_dense = big && dense_ok && (!selective || adapt_dense);
    if (want_dense) want_tc = false;
    if (const char* e = getenv("SB200_VIS_KERNEL")) {
      if (!strcmp(e, "simt")) { want_tc = false; want_dense = false; }
      else if (!strcmp(e, "tc")) { want_tc = work > 0; want_dense = false; }
      else if (!strcmp(e, "dense")) { want_dense = work > 0 && dense_ok; want_tc = work > 0 && !want_dense; }
    }
  }
  const int vote_cap = want_dense ? kDenseVoteCap : sb::kVoteVisCap;
  sb::Params Pf = P;
  Pf.vote_vis_cap = vote_cap;
  sb::SceneReq* hreq = q.h_req.as<sb::SceneReq>();
  for (int s = 0; s < n_scenes; ++s) {
    sb::SceneReq& r = hreq[s];
    r.slot = last_req_slots[s];
    r.m = m_of[s];
    r.det_base = det_offsets[s];
    r.epoch = epoch[r.slot] + 1;   // EpochDb::next_epoch, src/trackers/epoch_db.rs:35-49 (committed below)
    r.scene_id = scene_ids[s];
    r.pos_lbase = (int)posl_total;
    r.pos_lcap = (int)std::min<long long>((long long)r.m * 32, (long long)sb::kVotePosCap * 2);
    posl_total += r.pos_lcap;
    r.vis_lbase = (int)visl_total;
    r.vis_lcap = P.is_visual ? (int)std::min<long long>((long long)r.m * 64, (long long)vote_cap * 4) : 0;
    visl_total += r.vis_lcap;
  }
  // ---- frame buffers (sized once from the capacity hints when given, so steady-state frames never reallocate)
  const long long hint_dets = (long long)std::max(opts.max_scenes_hint, n_scenes) * opts.max_dets_per_scene_hint;
  const size_t T = (size_t)std::max<long long>(std::max(total, 1), hint_dets);
  const int hint_tracks = opts.max_tracks_per_scene_hint > 0 ? track_cap : 0;   // the store's rows per scene: hint + pipeline room
  const long long hint_cols = (long long)std::max(opts.max_scenes_hint, n_scenes) * (((long long)hint_tracks * K + 127) / 128 * 128);
  const long long pos_ub = pos_total;
  {
    const long long hint_pos = hint_dets * hint_tracks;
    pos_total = std::max(pos_total, hint_pos);
    if (P.is_visual) vis_total = std::max(vis_total, hint_pos * K);
  }
  // candidate-side buffers: the set of this frame (the other one may still be read by the frame in front of it)
  const int cset = (int)(frame_seq & 1);
  DBuf &f_cbox = f_cbox2[cset], &f_cradius = f_cradius2[cset], &f_cconf = f_cconf2[cset], &f_cvert = f_cvert2[cset],
       &f_cflags = f_cflags2[cset], &f_cnorm2 = f_cnorm22[cset], &f_cbf16 = f_cbf162[cset], &f_decided = f_decided2[cset];
  if ((rc = ENS(f_cbox, T * 24)) || (rc = ENS(f_cradius, T * 4)) || (rc = ENS(f_cconf, T * 4)) ||
      (rc = ENS(f_winner, T * 4)) || (rc = ENS(f_cvt, T)) || (rc = ENS(f_scenes, sizeof(sb::SceneDesc) * n_scenes)) ||
      (rc = ENS(f_newcount, 4 * (size_t)n_scenes)) || (rc = ENS(f_dyn, sizeof(sb::FrameDyn))) ||
      (rc = ENS(f_apprank, T * 8)) || (rc = ENS(f_appmeta, 16 * (size_t)n_scenes)) ||
      (rc = ENS(f_pos, std::max<size_t>(4, (size_t)pos_total * 4))))
    return rc;
  if (P.positional_kind == SB200_POS_IOU && (rc = ENS(f_cvert, T * 64))) return rc;
  if (P.is_visual) {
    if ((rc = ENS(f_cflags, T)) || (rc = ENS(f_cnorm2, T * 4)) || (rc = ENS(f_featdst, T * 4)) ||
        (rc = ENS(f_vis, std::max<size_t>(4, (size_t)vis_total * 4))) || (rc = ENS(f_scene_max, 4 * (size_t)n_scenes)))
      return rc;
  }
  sb::TcArgs tc;
  memset(&tc, 0, sizeof(tc));
  tc.num_sms = num_sms;
  int mstep = 0, cstep = 256;
  const long long visl_alloc = std::max(visl_total, hint_dets * 64);
  if (want_tc || want_dense) {
    tc.use_tc = true;
    tc.dense = want_dense;
    {
      // SB200_SCREEN = single | multicast | pair (default): CTA organisation of the screen kernel (the dense kernel: pairs)
      const char* e = getenv("SB200_SCREEN");
      tc.cluster2 = want_dense || (!(e && !strcmp(e, "single")) && getenv("SB200_SCREEN_SINGLE") == nullptr);
      tc.pair = want_dense || (tc.cluster2 && !(e && !strcmp(e, "multicast")));
    }
    mstep = tc.cluster2 ? 256 : 128;   // a cluster covers two 128-row candidate tiles
    if (want_dense) cstep = (256 / K) * K;   // column tiles end at block boundaries: a track's observations stay together
    long long tiles_ub = 0, slabs_ub = 0, ws_ub = 0, blk_ub = 0;
    for (int s = 0; s < n_scenes; ++s) {
      const long long ct = ((long long)nb_ub[s] * K + cstep - 1) / cstep;
      tiles_ub += (long long)((m_of[s] + mstep - 1) / mstep) * ct;
      slabs_ub += ct;
      ws_ub += ct * (cstep / K) * ((m_of[s] + 127) / 128 * 128);   // blocks padded to whole column tiles
      blk_ub += ct * (cstep / K);
    }
    tc.n_tiles = (int)tiles_ub;   // upper bound: the list and its length are built on the device
    // the list is sized from the hints like the other frame buffers (the bound grows with the frames in flight: a list
    // sized for this frame alone would be outgrown, and wait for the device, again and again)
    long long tiles_alloc = tiles_ub;
    {
      const long long hs = std::max(opts.max_scenes_hint, n_scenes), hd = std::max(opts.max_dets_per_scene_hint, max_m);
      const long long ht = std::max(hint_tracks, max_nb);
      tiles_alloc = std::max(tiles_alloc, hs * ((hd + mstep - 1) / mstep) * ((ht * K + cstep - 1) / cstep + 1));
      if (f_tiles.bytes > 0 && sizeof(sb::TcTile) * (size_t)tiles_alloc > f_tiles.bytes) tiles_alloc += tiles_alloc / 2;
    }
    if ((rc = ENS(f_cbf16, T * P.d8 * 2)) || (rc = ENS(f_tiles, sizeof(sb::TcTile) * (size_t)std::max<long long>(1, tiles_alloc))) ||
        (rc = ENS(f_rowmeta, sizeof(sb::VisRowMeta) * (T + 256))))
      return rc;
    tc.rowmeta = f_rowmeta.as<sb::VisRowMeta>();
    tc.max_rows = max_nb * K;
    tc.d_tiles = f_tiles.as<sb::TcTile>();
    tc.d_n_tiles = &f_dyn.as<sb::FrameDyn>()->n_tiles;
    tc.a_rows = total;
    tc.b_rows = (long long)scene_cap * track_cap * K;
    if (!want_dense) {
      if ((rc = ENS(f_colmeta, sizeof(sb::VisColMeta) * (size_t)(std::max(col_total, hint_cols) + 256))) ||
          (rc = ENS(f_colb, 4 * (size_t)(std::max(col_total, hint_cols) + 256))) ||
          (rc = ENS(f_colvalid, (size_t)(std::max(col_total, hint_cols) + 256) / 8 + 16)) ||
          (rc = ENS(f_colgeo, sizeof(sb::VisColGeo) * (size_t)std::max<long long>(1, P.n_constraints > 0 ? std::max(col_total, hint_cols) : 1))))
        return rc;
      tc.colmeta = f_colmeta.as<sb::VisColMeta>();
      tc.colgeo = f_colgeo.as<sb::VisColGeo>();
      tc.colb = f_colb.as<float>();
      tc.colvalid = f_colvalid.as<unsigned int>();
      tc.total_cols = (int)col_total;
    } else {
      // sized from the hints like the other frame buffers, so steady-state frames never reallocate
      const long long hs = std::max(opts.max_scenes_hint, n_scenes), ht = hint_tracks;
      const long long hd = std::max(opts.max_dets_per_scene_hint, 0);
      ws_ub = std::max(ws_ub, hs * (ht + 256) * ((hd + 127) / 128 * 128));
      blk_ub = std::max(blk_ub, hs * (ht + 256));
      slabs_ub = std::max(slabs_ub, hs * ((ht * K + cstep - 1) / cstep + 1));
      if ((rc = ENS(f_drowb, 4 * 5 * T)) || (rc = ENS(f_dcolb, 4 * (size_t)std::max<long long>(1, blk_ub))) ||
          (rc = ENS(f_slabk, 4 * 256 * (size_t)std::max<long long>(1, slabs_ub))))
        return rc;
      if ((rc = ENS(f_ws, 4 * (size_t)std::max<long long>(1, ws_ub))) || (rc = ENS(f_tmeta, sizeof(sb::DenseTrackMeta) * (size_t)std::max<long long>(1, blk_ub))) ||
          (rc = ENS(f_rowinfo, 8 * (size_t)std::max<long long>(1, blk_ub * K))) || (rc = ENS(f_slabc, 4 * 256 * (size_t)std::max<long long>(1, slabs_ub))) ||
          (rc = ENS(f_slabm, 4 * 256 * (size_t)std::max<long long>(1, slabs_ub))) || (rc = ENS(f_slabmask, 2 * 32 * (size_t)std::max<long long>(1, slabs_ub))) ||
          (rc = ENS(f_dscene, 4 * 6 * (size_t)n_scenes + 128)) || (rc = ENS(f_maxc, sizeof(sb::VisPair) * (size_t)std::max<long long>(1, visl_alloc))) ||
          (rc = ENS(f_maxcval, 4 * (size_t)std::max<long long>(1, visl_alloc))))
        return rc;
      tc.cstep = cstep;
      tc.max_blocks = max_nb;
      tc.n_slabs_ub = (int)slabs_ub;
      tc.ws = f_ws.p;
      tc.d_rowb = f_drowb.as<unsigned int>();
      tc.d_colb = f_dcolb.as<unsigned int>();
      tc.blk_ub = blk_ub;
      tc.slab_ktf = f_slabk.as<float>();
      tc.tmeta = f_tmeta.as<sb::DenseTrackMeta>();
      tc.rowinfo = f_rowinfo.as<int2>();
      tc.slab_colc = f_slabc.as<float>();
      tc.slab_cmax = f_slabm.as<float>();
      tc.slab_vmask = f_slabmask.as<unsigned int>();
      tc.slab_bmask = tc.slab_vmask + (size_t)slabs_ub * 8;
      // per-scene scalars: maxc_cnt | maxc_next | dense_bad | zeros (ints, zeroed together), then l0 | cmax (floats)
      tc.maxc_cnt = f_dscene.as<int>();
      tc.maxc_next = tc.maxc_cnt + n_scenes;
      tc.dense_bad = tc.maxc_cnt + 2 * n_scenes;
      tc.zeros = tc.maxc_cnt + 3 * n_scenes;
      tc.scene_l0 = reinterpret_cast<float*>(tc.maxc_cnt + 4 * n_scenes);
      tc.scene_cmax = tc.scene_l0 + n_scenes;
      tc.dbg_counts = reinterpret_cast<int*>(tc.scene_cmax + n_scenes) + 4;
      tc.maxc = f_maxc.as<sb::VisPair>();
      tc.maxc_val = f_maxcval.as<float>();
    }
  }
  sb::Frame f;
  memset(&f, 0, sizeof(f));
  f.total = total;
  f.pos_total = pos_ub;
  f.dyn = f_dyn.as<sb::FrameDyn>();
  f.id_counter = b_idc.as<unsigned long long>();
  f.id_add = P.is_batch ? (long long)total : -1;
  f.dense_bad = tc.dense ? tc.dense_bad : nullptr;
  // dense positional matrices for every scene only on request (SB200_FULL_COSTS / SB200_NO_FORK: parity of sb200_last_costs)
  f.pos_dense_all = getenv("SB200_FULL_COSTS") != nullptr || getenv("SB200_NO_FORK") != nullptr;
  bool prefetched = false;
  Staging* sin = nullptr;
  // inputs
  if (device_io) {
    f.in_boxes = boxes; f.in_feat = features; f.in_hasf = has_feature; f.in_quality = quality;
    f.in_custom = reinterpret_cast<const long long*>(custom_ids); f.in_own = own_area;
  } else {
    // device staging: a set already filled by sb200_prefetch_inputs() for exactly these host buffers (all six columns)
    // is used as is; otherwise the H2D copies are issued below
    const void* key[6] = {boxes, features, features ? has_feature : nullptr, quality, custom_ids, own_area};
    int use = -1;
    for (int k = 0; k < 2; ++k)
      if (stg[k].pending && stg[k].total == total && memcmp(stg[k].key, key, sizeof(key)) == 0) use = k;
    if (use >= 0) {
      prefetched = true;
      stg[use].pending = false;
      CU(cudaStreamWaitEvent(stream, stg[use].ev, 0));
    } else {
      use = stg[0].pending ? 1 : (stg[1].pending ? 0 : 1 - stg_last);
      stg[use].pending = false;
    }
    stg_last = use;
    Staging& S = stg[use];
    sin = &S;
    if (prefetched && (S.boxes.bytes < T * 24 || (features && S.feat.bytes < T * (size_t)P.feature_dim * 4))) {
      // the filled set is smaller than this frame's sizing rule (hints changed?): re-copy instead of reallocating
      prefetched = false;
    }
    if ((rc = ENS(S.boxes, T * 24))) return rc;
    f.in_boxes = S.boxes.as<float>();
    if (features && total > 0) {
      if ((rc = ENS(S.feat, T * (size_t)P.feature_dim * 4))) return rc;
      f.in_feat = S.feat.as<float>();
      if (has_feature) {
        if ((rc = ENS(S.hasf, T))) return rc;
        f.in_hasf = S.hasf.as<unsigned char>();
      }
    }
    if (quality && total > 0) { if ((rc = ENS(S.quality, T * 4))) return rc; f.in_quality = S.quality.as<float>(); }
    if (custom_ids && total > 0) { if ((rc = ENS(S.custom, T * 8))) return rc; f.in_custom = S.custom.as<long long>(); }
    if (own_area && total > 0) { if ((rc = ENS(S.own, T * 4))) return rc; f.in_own = S.own.as<float>(); }
  }
  f.c_box = f_cbox.as<float>(); f.c_radius = f_cradius.as<float>(); f.c_conf = f_cconf.as<float>();
  f.c_vert = f_cvert.as<double>(); f.c_flags = f_cflags.as<unsigned char>(); f.c_norm2 = f_cnorm2.as<float>();
  f.winner = f_winner.as<int>(); f.c_vt = f_cvt.as<unsigned char>(); f.pos = f_pos.as<float>(); f.vis = f_vis.as<float>();
  f.scenes = f_scenes.as<sb::SceneDesc>(); f.new_count = f_newcount.as<int>();
  f.new_count_all = f.new_count;
  f.feat_dst = P.is_visual ? f_featdst.as<int>() : nullptr;
  f.app_rank = f_apprank.as<int2>(); f.app_meta = f_appmeta.as<int4>();
  if ((rc = ENS(f_frameout, sizeof(int) * 3 * (size_t)n_scenes))) return rc;
  f.frame_out = f_frameout.as<int>();
  f.c_bf16 = tc.use_tc ? f_cbf16.p : nullptr; f.scene_max = f_scene_max.as<unsigned int>();
  // sparse entry lists + per-scene counters (pos_cnt | vis_cnt | scene_mode | vis_mode | refine_next | dense_cnt | status),
  // zeroed every frame by frame_setup_kernel
  const size_t n_counters = 6 * (size_t)n_scenes + 4;
  if ((rc = ENS(f_poslist, sizeof(sb::PosEntry) * (size_t)std::max<long long>(1, std::max(posl_total, hint_dets * 32)))) ||
      (rc = ENS(f_counters, sizeof(int) * n_counters)))
    return rc;
  if (P.is_visual && ((rc = ENS(f_pairs, sizeof(sb::VisPair) * (size_t)std::max<long long>(1, visl_alloc))) ||
                      (rc = ENS(f_visval, sizeof(float) * (size_t)std::max<long long>(1, visl_alloc)))))
    return rc;
  f.pos_list = f_poslist.as<sb::PosEntry>();
  f.pos_cnt = f_counters.as<int>();
  f.vis_cnt = f.pos_cnt + n_scenes;
  f.scene_mode = f.pos_cnt + 2 * n_scenes;
  f.vis_mode = f.pos_cnt + 3 * n_scenes;
  f.refine_next = f.pos_cnt + 4 * n_scenes;
  f.status = f.pos_cnt + 5 * n_scenes;
  f.dense_cnt = f.pos_cnt + 6 * n_scenes;
  // SB200_POS_GQ=1 (experiment): gated positional pairs of the frame in one queue, evaluated by a kernel of their own
  if (getenv("SB200_POS_GQ") != nullptr) {
    size_t gq_cap = (size_t)std::min<long long>(std::max<long long>(1, std::max<long long>(total, hint_dets) * 24), 1ll << 26);
    if (const char* e = getenv("SB200_POS_GQ_CAP")) gq_cap = (size_t)std::max(1, atoi(e));   // tests: force the overflow path
    if ((rc = ENS(f_posgq, 8 * gq_cap))) return rc;
    f.pos_gq = f_posgq.as<int2>();
    f.pos_gq_cnt = f.pos_cnt + 6 * n_scenes + 1;
    f.pos_gq_cap = (int)gq_cap;
  }
  f.vis_pairs = f_pairs.as<sb::VisPair>();
  f.vis_val = f_visval.as<float>();
  // outputs
  sb200_predict_out o{};
  if (out) o = *out;
  if (device_io) {
    f.o_ids = reinterpret_cast<unsigned long long*>(o.ids); f.o_epochs = o.epochs; f.o_lengths = o.lengths;
    f.o_vt = o.voting_types; f.o_pred = o.predicted_boxes; f.o_obs = o.observed_boxes;
  } else {
    if (o.ids) { if ((rc = ENS(o_ids, T * 8))) return rc; f.o_ids = o_ids.as<unsigned long long>(); }
    if (o.epochs) { if ((rc = ENS(o_epochs, T * 4))) return rc; f.o_epochs = o_epochs.as<unsigned int>(); }
    if (o.lengths) { if ((rc = ENS(o_lengths, T * 4))) return rc; f.o_lengths = o_lengths.as<unsigned int>(); }
    if (o.voting_types) { if ((rc = ENS(o_vt, T))) return rc; f.o_vt = o_vt.as<unsigned char>(); }
    if (o.predicted_boxes) { if ((rc = ENS(o_pred, T * 24))) return rc; f.o_pred = o_pred.as<float>(); }
    if (o.observed_boxes) { if ((rc = ENS(o_obs, T * 24))) return rc; f.o_obs = o_obs.as<float>(); }
  }
  // Visual trackers on the tensor-core path evaluate the positional metric lazily: VisualVoting only consults it for
  // candidates the visual BestFit pass left undecided, against tracks that pass did not claim, so the order is
  // screen -> refine -> BestFit pre-pass (masks) -> culled scan of what is still open -> full voting.  The dense None
  // fill of the positional matrices runs on a side stream next to the screen.  SB200_FULL_COSTS=1 (every pair is
  // evaluated, sb200_last_costs is complete) and SB200_NO_FORK=1 keep the plain order.
  const bool full_costs = getenv("SB200_FULL_COSTS") != nullptr || getenv("SB200_NO_FORK") != nullptr;
  const bool fork = P.is_visual && tc.use_tc && tc.n_tiles > 0 && !full_costs;
  if (fork) {
    if ((rc = ENS(f_decided, T)) || (rc = ENS(f_excl, (size_t)scene_cap * track_cap + 16)) || (rc = ENS(f_prewin, T * 4))) return rc;
    f.decided = f_decided.as<unsigned char>();
    f.excl = f_excl.as<unsigned char>();
    f.pre_winner = f_prewin.as<int>();
  }
  const bool derive_own = P.is_visual && P.use_own_area && f.in_own == nullptr && total > 0;
  if (derive_own && ((rc = ENS(f_own, T * 4)) || (rc = ENS(f_ownovf, 16 + T * sizeof(int2))))) return rc;

  // ======================================================================== nothing below can fail for capacity reasons:
  // the request is committed (epochs, ring slot, bounds of the frames in flight)
  for (int s = 0; s < n_scenes; ++s) epoch[last_req_slots[s]] += 1;
  if (features != nullptr && total > 0) seen_features = true;
  frame_seq += 1;
  q.active = true;
  q.n_scenes = n_scenes;
  q.total = total;
  q.slots.assign(last_req_slots.begin(), last_req_slots.end());
  q.m = m_of;
  q.live_ub = live_ub;
  q.tc_timed = false;
  q.pos_forked = false;
  q.mode = tc.dense ? 2 : (tc.use_tc ? 1 : 0);
  for (int s = 0; s < n_scenes; ++s) pending_add[last_req_slots[s]] += m_of[s];
  inflight_live_ub += live_ub;
  pend_count += 1;
  last_n_scenes = n_scenes;
  // a CUDA failure while the frame is being enqueued takes it out of the ring again (the context is lost anyway)
  struct Rollback {
    sb200_tracker* t; Pending* q; bool armed;
    ~Rollback() {
      if (!armed) return;
      for (int s = 0; s < q->n_scenes; ++s) t->pending_add[q->slots[s]] -= q->m[s];
      t->inflight_live_ub -= q->live_ub;
      q->active = false;
      t->pend_count -= 1;
    }
  } rollback{this, &q, true};

  const double ms_setup = since(t_begin);
  if (has_user_stream) {   // everything the caller's stream holds now comes first
    if (!ev_user_in) { CU(cudaEventCreateWithFlags(&ev_user_in, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&ev_user_out, cudaEventDisableTiming)); }
    CU(cudaEventRecord(ev_user_in, user_stream));
    CU(cudaStreamWaitEvent(stream, ev_user_in, 0));
  }
  if (!device_io && !prefetched && total > 0) {
    const size_t n = (size_t)total;
    CU(cudaMemcpyAsync(sin->boxes.p, boxes, n * 24, cudaMemcpyHostToDevice, stream));
    if (f.in_feat) {
      CU(cudaMemcpyAsync(sin->feat.p, features, n * (size_t)P.feature_dim * 4, cudaMemcpyHostToDevice, stream));
      if (f.in_hasf) CU(cudaMemcpyAsync(sin->hasf.p, has_feature, n, cudaMemcpyHostToDevice, stream));
    }
    if (f.in_quality) CU(cudaMemcpyAsync(sin->quality.p, quality, n * 4, cudaMemcpyHostToDevice, stream));
    if (f.in_custom) CU(cudaMemcpyAsync(sin->custom.p, custom_ids, n * 8, cudaMemcpyHostToDevice, stream));
    if (f.in_own) CU(cudaMemcpyAsync(sin->own.p, own_area, n * 4, cudaMemcpyHostToDevice, stream));
  }
  // scene descriptors, tile list, frame scalars; list counters and status words zeroed
  if (!pos_stream && !side_off) {
    static const bool off = [] { const char* e = getenv("SB200_SIDE_STREAM"); return e && e[0] == '0'; }();
    side_off = off;
    if (!side_off) {
      int lo = 0, hi = 0;
      CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CU(cudaStreamCreateWithPriority(&pos_stream, cudaStreamNonBlocking, hi));
      for (auto& e : ev_fork) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
  }
  const bool side_ok = pos_stream != nullptr;
  // The frame's tables come from one small CTA: with nothing else to wait for (no own-area derivation, which reads them) it
  // runs beside the candidate preparation, on the side stream, and the main stream joins before the first kernel that reads them.
  if (!prep_stream && !prep_off) {
    static const bool off = [] { const char* e = getenv("SB200_PREP_AHEAD"); return e && e[0] == '0'; }();
    prep_off = off;
    if (!prep_off) {
      int lo = 0, hi = 0;
      CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CU(cudaStreamCreateWithPriority(&prep_stream, cudaStreamNonBlocking, lo));
      CU(cudaEventCreateWithFlags(&ev_prep_done, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&ev_inputs, cudaEventDisableTiming));
    }
  }
  const bool prep_ahead = prep_stream != nullptr && !derive_own && total > 0;
  // (measured: keeping the tables on the work stream when the preparation runs ahead saves nothing -- 0.873 vs 0.875 ms -- and
  // moves the next frame's preparation squarely under the screen kernel, 0.260 vs 0.248 ms; the side stream stays)
  const bool side_setup = side_ok && !derive_own && total > 0;
  cudaStream_t s_setup = stream;
  if (side_setup) {
    CU(cudaEventRecord(ev_fork[0], stream));
    CU(cudaStreamWaitEvent(pos_stream, ev_fork[0], 0));
    s_setup = pos_stream;
  }
  sb::launch_frame_setup(P, ts, f, reinterpret_cast<const sb::SceneReq*>(q.h_req.dp), n_scenes, b_ntracks.as<int>(), mstep,
                         cstep, tc.dense, f_tiles.as<sb::TcTile>(), f_dyn.as<sb::FrameDyn>(), f_counters.as<int>(), (int)n_counters, s_setup);
  tc.max_init_done = 1;   // frame_setup resets scene_max
  if (side_setup && Pf.is_visual && tc.use_tc && !tc.dense && tc.n_tiles > 0 && max_m > 0 && max_n > 0 && f.in_feat) {
    // the screen's column metadata reads the tables and the store only: it follows the setup on the side stream
    sb::launch_vis_colmeta(Pf, ts, f, n_scenes, max_n, tc, pos_stream);
    tc.colmeta_done = 1;
  }
  if (side_setup) CU(cudaEventRecord(ev_join, pos_stream));
  CU(cudaEventRecord(q.ev[0], stream));
  if (derive_own) {
    // visual_sort/simple_api.rs:110-127: with an own-area threshold and no shares supplied by the caller, the shares come
    // from the scene's observation boxes (exclusively_owned_areas_normalized_shares)
    sb::launch_own_area(f, n_scenes, max_m, f.in_boxes, f_own.as<float>(), f_ownovf.as<int>(),
                        reinterpret_cast<int2*>(f_ownovf.as<char>() + 16), stream);
    f.in_own = f_own.as<float>();
  }
  // Candidate preparation needs the request only (boxes, features): it runs on its own stream as soon as the inputs are there
  // and the frame that read this set of candidate buffers (two frames back) has ended -- in the steady state under the
  // tensor-core kernel of the frame in front.  (With derived own-area shares it needs the frame tables: main stream.)
  if (prep_ahead) {
    if (has_user_stream) CU(cudaStreamWaitEvent(prep_stream, ev_user_in, 0));
    if (!device_io) {   // staged inputs: the prefetch copy, or the copies issued above on the work stream
      if (prefetched) CU(cudaStreamWaitEvent(prep_stream, sin->ev, 0));
      else { CU(cudaEventRecord(ev_inputs, stream)); CU(cudaStreamWaitEvent(prep_stream, ev_inputs, 0)); }
    }
    if (set_busy[cset]) CU(cudaStreamWaitEvent(prep_stream, ev_set_free[cset], 0));
    // ... and not before the frame in front has left its cost kernels: the preparation is HBM traffic; under that frame's
    // tensor-core screen it costs the screen 3-6 % (0.248-0.26 ms instead of 0.243), under its voting / apply / feature store it
    // costs those ~40 us.  Measured at cfg5: 0.897 ms/step with the screen at 0.717 of the BF16 peak here, against 0.875-0.88
    // ms/step with the screen at 0.67-0.70 for SB200_PREP_AFTER=none (as early as possible).  The default keeps the
    // tensor-core kernel undisturbed and the step time reproducible; a deployment that only counts frames sets "none".
    static const bool after_cost = [] { const char* e = getenv("SB200_PREP_AFTER"); return !(e && !strcmp(e, "none")); }();
    if (after_cost && cost_done_valid) CU(cudaStreamWaitEvent(prep_stream, ev_cost_done, 0));
    sb::launch_prep(Pf, f, n_scenes, max_m, prep_stream);
    CU(cudaEventRecord(ev_prep_done, prep_stream));
    CU(cudaStreamWaitEvent(stream, ev_prep_done, 0));
  } else {
    sb::launch_prep(Pf, f, n_scenes, max_m, stream);
  }
  if (side_setup) CU(cudaStreamWaitEvent(stream, ev_join, 0));
  CU(cudaEventRecord(q.ev[1], stream));
  sb::TcArgs tcc = tc;
  if (tc.use_tc && tc.n_tiles > 0) { tcc.ev_screen0 = q.ev_k[0]; tcc.ev_screen1 = q.ev_k[1]; tcc.ev_refine1 = q.ev_k[2]; q.tc_timed = true; }
  if (!fork) {
    sb::launch_pos_cost(Pf, ts, f, n_scenes, max_m, max_n, stream);
    CU(cudaEventRecord(q.ev[2], stream));
    int vr0 = sb::launch_vis_cost(Pf, ts, f, n_scenes, max_m, max_n, tcc, stream);
    if (vr0 != 0) return fail(SB200_ERR_CUDA, "visual cost launch failed (%d)", vr0);
    // scenes whose entry list overflowed vote on the dense matrix: it is filled and scanned for them alone, now
    if (!f.pos_dense_all) sb::launch_pos_scan_lazy(Pf, ts, f, n_scenes, max_m, max_n, /*pass=*/1, stream);
  } else {
    CU(cudaEventRecord(q.ev[2], stream));
    {
      int vr0 = sb::launch_vis_cost_a(Pf, ts, f, n_scenes, max_m, max_n, tcc, stream);   // metadata, screen, vis_mode, refinement
      if (vr0 != 0) return fail(SB200_ERR_CUDA, "visual cost launch failed (%d)", vr0);
      vr0 = sb::launch_vote_masks(Pf, ts, f, n_scenes, max_m, max_n, stream);            // who is still open positionally
      if (vr0 != 0) return fail(SB200_ERR_CUDA, "voting launch failed: %s", cudaGetErrorString((cudaError_t)vr0));
    }
    CU(cudaEventRecord(q.ev_pos[0], stream));
    sb::launch_pos_scan_lazy(Pf, ts, f, n_scenes, max_m, max_n, /*pass=*/0, stream);
    CU(cudaEventRecord(q.ev_pos[1], stream));
    int vr1 = sb::launch_vis_cost_b(Pf, ts, f, n_scenes, max_m, max_n, tcc, stream);     // final scene mode, dense fallbacks
    if (vr1 != 0) return fail(SB200_ERR_CUDA, "visual cost launch failed (%d)", vr1);
    sb::launch_pos_scan_lazy(Pf, ts, f, n_scenes, max_m, max_n, /*pass=*/1, stream);     // scenes that fell back to dense voting
    q.pos_forked = true;
  }
  CU(cudaEventRecord(q.ev[3], stream));
  if (prep_stream) {
    if (!ev_cost_done) CU(cudaEventCreateWithFlags(&ev_cost_done, cudaEventDisableTiming));
    CU(cudaEventRecord(ev_cost_done, stream));
    cost_done_valid = true;
  }
  int vr = sb::launch_voting(Pf, ts, f, n_scenes, max_m, max_n, stream);
  if (vr != 0) return fail(SB200_ERR_CUDA, "voting launch failed: %s", cudaGetErrorString((cudaError_t)vr));
  CU(cudaEventRecord(q.ev[4], stream));
  sb::launch_apply(Pf, ts, f, n_scenes, max_m, 0ull, b_ntracks.as<int>(), stream);
  // The sweep (latency-bound, one CTA per scene) and the feature store (HBM-bound) touch disjoint arrays -- a track's feature
  // block is not moved by the compaction -- so the sweep runs on the side stream beside the store; the frame ends at the join.
  const bool side_sweep = side_ok && Pf.is_visual && f.in_feat && total > 0;
  if (side_sweep) {
    CU(cudaEventRecord(ev_fork[1], stream));
    CU(cudaStreamWaitEvent(pos_stream, ev_fork[1], 0));
    sb::launch_frame_sweep(Pf, ts, f, n_scenes, b_ntracks.as<int>(), wb, pos_stream);
    CU(cudaEventRecord(ev_join, pos_stream));
    sb::launch_feat_store(Pf, ts, f, stream);
    CU(cudaStreamWaitEvent(stream, ev_join, 0));
  } else {
    sb::launch_feat_store(Pf, ts, f, stream);
    sb::launch_frame_sweep(Pf, ts, f, n_scenes, b_ntracks.as<int>(), wb, stream);
  }
  CU(cudaEventRecord(q.ev[5], stream));
  CU(cudaGetLastError());

  // results back
  if (!device_io && total > 0) {
    if (o.ids) CU(cudaMemcpyAsync(o.ids, f.o_ids, (size_t)total * 8, cudaMemcpyDeviceToHost, stream));
    if (o.epochs) CU(cudaMemcpyAsync(o.epochs, f.o_epochs, (size_t)total * 4, cudaMemcpyDeviceToHost, stream));
    if (o.lengths) CU(cudaMemcpyAsync(o.lengths, f.o_lengths, (size_t)total * 4, cudaMemcpyDeviceToHost, stream));
    if (o.voting_types) CU(cudaMemcpyAsync(o.voting_types, f.o_vt, (size_t)total, cudaMemcpyDeviceToHost, stream));
    if (o.predicted_boxes) CU(cudaMemcpyAsync(o.predicted_boxes, f.o_pred, (size_t)total * 24, cudaMemcpyDeviceToHost, stream));
    if (o.observed_boxes) CU(cudaMemcpyAsync(o.observed_boxes, f.o_obs, (size_t)total * 24, cudaMemcpyDeviceToHost, stream));
  }
  {
    char* ho = reinterpret_cast<char*>(q.h_out.p);
    CU(cudaMemcpyAsync(ho, f.frame_out, 12 * (size_t)n_scenes, cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(ho + 12 * (size_t)n_scenes, f.status, 4 * (size_t)n_scenes, cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(ho + dyn_offset(n_scenes), f_dyn.p, sizeof(sb::FrameDyn), cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(ho + dyn_offset(n_scenes) + sizeof(sb::FrameDyn), f.dense_cnt, 4, cudaMemcpyDeviceToHost, stream));
  }
  static const bool trace_dense = trace && atoi(getenv("SB200_TRACE")) >= 2;   // synchronises: level 2 only
  if (trace_dense && tc.dense && tc.n_tiles > 0) {
    int hc[8] = {0};
    int vc = 0;
    CU(cudaMemcpyAsync(hc, tc.dbg_counts, sizeof(hc), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    std::vector<int> vcnt((size_t)n_scenes);
    CU(cudaMemcpy(vcnt.data(), f.vis_cnt, 4 * (size_t)n_scenes, cudaMemcpyDeviceToHost));
    long long tot = 0; int mx = 0;
    for (int v : vcnt) { tot += v; mx = std::max(mx, v); }
    (void)vc;
    fprintf(stderr, "[sb200] dense frame: fallback reasons threshold %d, max-list overflow %d, no maximum %d; max candidates %d; pairs to refine %lld (largest scene %d)\n",
            hc[0], hc[1], hc[2], hc[3], tot, mx);
  }
  CU(cudaEventRecord(q.done, stream));
  if (!ev_set_free[cset]) CU(cudaEventCreateWithFlags(&ev_set_free[cset], cudaEventDisableTiming));
  CU(cudaEventRecord(ev_set_free[cset], stream));   // this frame's candidate buffers may be rewritten after this point
  set_busy[cset] = true;
  if (has_user_stream && join_per_call) {   // the caller's stream continues after the frame
    CU(cudaEventRecord(ev_user_out, stream));
    CU(cudaStreamWaitEvent(user_stream, ev_user_out, 0));
  }
  rollback.armed = false;
  if (sin) { CU(cudaEventRecord(sin->ev_read, stream)); sin->read_pending = true; }
  const double ms_launch = since(t_begin);
  if (wait) rc = drain();
  host_ms_total += since(t_begin);
  host_calls += 1;
  if (trace && prefetched && sin && wait) {
    float cms = 0.0f;
    if (cudaEventElapsedTime(&cms, sin->ev0, sin->ev) == cudaSuccess)
      fprintf(stderr, "[sb200] prefetch copy of this frame took %.3f ms on the copy stream\n", cms);
  }
  if (trace) fprintf(stderr, "[sb200] predict: setup %.3f ms, launched at %.3f ms, returned at %.3f ms (total dets %d, %d in flight)\n", ms_setup, ms_launch, since(t_begin), total, pend_count);
  return rc;
}

#undef ENS
// =============================================================================================== C ABI
extern "C" {

const char* sb200_last_error(void) { return g_err.c_str(); }
void sb200__set_error(const char* msg) { g_err = msg ? msg : ""; }  // used by ops.cu / nms

int sb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void sb200_options_default(sb200_options* o) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->kind = SB200_KIND_SORT;
  o->positional_kind = SB200_POS_MAHA;
  o->iou_threshold = 0.3f;
  o->min_confidence = 0.05f;
  o->max_idle_epochs = 5;
  o->history_length = 1;
  o->kalman_position_weight = 1.0f / 20.0f;
  o->kalman_velocity_weight = 1.0f / 160.0f;
  o->visual_kind = SB200_VIS_EUCLIDEAN;
  o->visual_threshold = 3.402823466e+38f;
  o->visual_max_observations = 5;
  o->visual_min_votes = 1;
  o->visual_minimal_track_length = 3;
}

int sb200_tracker_create(const sb200_options* opts, sb200_tracker** out) {
  if (!opts || !out) return fail(SB200_ERR_INVALID, "opts / out is NULL");
  *out = nullptr;
  int ndev = sb200_device_count();
  if (ndev <= 0) return fail(SB200_ERR_CUDA, "no CUDA device available (this library has no CPU execution path)");
  if (opts->device < 0 || opts->device >= ndev) return fail(SB200_ERR_INVALID, "device %d out of range (%d devices)", opts->device, ndev);
  sb::Params P;
  int rc = make_params(*opts, &P);
  if (rc) return rc;
  CU(cudaSetDevice(opts->device));
  sb200_tracker* t = new sb200_tracker();
  t->opts = *opts;
  t->P = P;
  t->device = opts->device;
  // history_length 0 means "unlimited" in the reference (sort.rs:166); the device keeps at most kMaxHist boxes per track
  t->hist_len = opts->history_length <= 0 ? sb::kMaxHist : std::min<int>(opts->history_length, sb::kMaxHist);
  {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, opts->device) == cudaSuccess && sms > 0) t->num_sms = sms;
  }
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  // numerically lower = more urgent: side stream (prio_hi) > work stream > candidate preparation (prio_lo)
  cudaError_t e = cudaStreamCreateWithPriority(&t->stream, cudaStreamNonBlocking, prio_hi < prio_lo ? prio_hi + 1 : prio_lo);
  if (e != cudaSuccess) { delete t; return fail(SB200_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e)); }
  for (auto& g : t->stg) {
    e = cudaEventCreate(&g.ev);
    if (e == cudaSuccess) e = cudaEventCreate(&g.ev0);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g.ev_read, cudaEventDisableTiming);
    if (e != cudaSuccess) { delete t; return fail(SB200_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(e)); }
  }
  if (opts->max_scenes_hint > 0 || opts->max_tracks_per_scene_hint > 0) {
    // + room for the tracks the frames in flight may add (see predict())
    rc = t->ensure_store(std::max(1, opts->max_scenes_hint),
                         std::max(64, opts->max_tracks_per_scene_hint + sb200_tracker::kDepth * std::max(0, opts->max_dets_per_scene_hint)));
    if (rc) { delete t; return rc; }
  }
  *out = t;
  return 0;
}

void sb200_tracker_destroy(sb200_tracker* t) { delete t; }

int sb200_tracker_set_stream(sb200_tracker* t, void* cuda_stream) {
  if (!t) return fail(SB200_ERR_INVALID, "tracker is NULL");
  CU(cudaSetDevice(t->device));
  int rc = t->drain();
  if (rc) return rc;
  CU(cudaStreamSynchronize(t->stream));
  // the tracker keeps its own work stream; the caller's stream is joined by events at every call (see `user_stream`)
  t->user_stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  t->has_user_stream = true;
  return 0;
}

int sb200_set_stream_join(sb200_tracker* t, int32_t per_call) {
  if (!t) return fail(SB200_ERR_INVALID, "tracker is NULL");
  t->join_per_call = per_call != 0;
  return 0;
}

int sb200_stream_join(sb200_tracker* t, void* cuda_stream) {
  if (!t) return fail(SB200_ERR_INVALID, "tracker is NULL");
  CU(cudaSetDevice(t->device));
  if (!t->ev_join_req) CU(cudaEventCreateWithFlags(&t->ev_join_req, cudaEventDisableTiming));
  CU(cudaEventRecord(t->ev_join_req, t->stream));
  CU(cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(cuda_stream), t->ev_join_req, 0));
  return 0;
}

int sb200_predict_batch(sb200_tracker* t, int32_t n_scenes, const uint64_t* scene_ids, const int32_t* det_offsets,
                        const float* boxes, const float* features, const uint8_t* has_feature, const float* quality,
                        const int64_t* custom_ids, const float* own_area, const sb200_predict_out* out) {
  if (!t) return fail(SB200_ERR_INVALID, "tracker is NULL");
  return t->predict(n_scenes, scene_ids, det_offsets, boxes, features, has_feature, quality, custom_ids, own_area, out, false, true);
}

int sb200_predict_batch_async(sb200_tracker* t, int32_t n_scenes, const uint64_t* scene_ids, const int32_t* det_offsets,
                              const float* boxes, const float* features, const uint8_t* has_feature, const float* quality,
                              const int64_t* custom_ids, const float* own_area, const sb200_predict_out* out) {
  if (!t) return fail(SB200_ERR_INVALID, "tracker is NULL");
  return t->predict(n_scenes, scene_ids, det_offsets, boxes, features, has_feature, quality, custom_ids, own_area, out, false, false);
}

int sb200_set_feature_dim(sb200_tracker* t, int32_t feature_dim) {
  if (!t || feature_dim <= 0) return fail(SB200_ERR_INVALID, "bad arguments");
  if (!t->P.is_visual) return fail(SB200_ERR_INVALID, "not a visual tracker");
  if (feature_dim == t->P.feature_dim) return 0;
  if (t->seen_features) return fail(SB200_ERR_INVALID, "features of dimension %d are already stored", t->P.feature_dim);
  CU(cudaSetDevice(t->device));
  int rc = t->drain();
  if (rc) return rc;
  CU(cudaStreamSynchronize(t->stream));
  // no track holds a feature yet (obs_hasf == 0 everywhere): the feature arena is simply re-created for the new row size
  t->P.feature_dim = feature_dim;
  t->P.d8 = (feature_dim + 7) / 8 * 8;
  t->opts.feature_dim = feature_dim;
  t->b_feat.release(); t->b_feat_bf16.release(); t->f_cbf162[0].release(); t->f_cbf162[1].release();
  t->ts.feat = nullptr; t->ts.feat_bf16 = nullptr;
  const size_t rows = (size_t)t->scene_cap * t->track_cap * t->P.max_obs;
  if (rows > 0) {
    if ((rc = t->b_feat.ensure(rows * t->P.d8 * 4)) || (rc = t->b_feat_bf16.ensure(rows * t->P.d8 * 2))) return rc;
    t->ts.feat = t->b_feat.as<float>();
    t->ts.feat_bf16 = t->b_feat_bf16.p;
  }
  for (auto& g : t->stg) g.feat.release();
  return 0;
}

int sb200_sync(sb200_tracker* t) {
  if (!t) return fail(SB200_ERR_INVALID, "tracker is NULL");
  CU(cudaSetDevice(t->device));
  return t->drain();
}

int sb200_frames_in_flight(sb200_tracker* t) {
  if (!t) return fail(SB200_ERR_INVALID, "tracker is NULL");
  CU(cudaSetDevice(t->device));
  t->poll();
  return t->pend_count;
}

int sb200_work_counters(sb200_tracker* t, uint64_t* out3 /* [4] */, double* ms7 /* [8] */) {
  if (!t) return fail(SB200_ERR_INVALID, "tracker is NULL");
  CU(cudaSetDevice(t->device));
  int rc = t->drain();
  if (rc) return rc;
  if (out3) { out3[0] = t->acc_units_mn; out3[1] = t->acc_units_rows; out3[2] = t->acc_frames; out3[3] = t->acc_dense_scenes; }
  if (ms7) {
    for (int i = 0; i < 5; ++i) ms7[i] = t->acc_stage_ms[i];
    ms7[5] = t->acc_kernel_ms[0]; ms7[6] = t->acc_kernel_ms[1]; ms7[7] = (double)t->acc_tc_frames;
  }
  return 0;
}

uint64_t sb200_launch_count(void) { return sb::launch_count(); }

int sb200_host_counters(sb200_tracker* t, double* out3 /* [3] */) {
  if (!t || !out3) return fail(SB200_ERR_INVALID, "bad arguments");
  out3[0] = (double)t->host_calls; out3[1] = t->host_ms_total; out3[2] = t->host_ms_blocked;
  return 0;
}

int sb200_prefetch_inputs(sb200_tracker* t, int32_t total, const float* boxes, const float* features,
                          const uint8_t* has_feature, const float* quality, const int64_t* custom_ids,
                          const float* own_area) {
  if (!t || total < 0 || (total > 0 && !boxes)) return fail(SB200_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(t->device));
  if (total == 0) return 0;
  if (!t->P.is_visual) { features = nullptr; has_feature = nullptr; quality = nullptr; own_area = nullptr; }
  if (!t->copy_stream) CU(cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking));
  // a free set: not holding an unconsumed prefetch; with nothing pending, the one the last predict did not read
  if (t->stg[0].pending && t->stg[1].pending)
    return fail(SB200_ERR_INVALID, "two prefetched requests are already waiting for their predict call");
  const int k = t->stg[0].pending ? 1 : (t->stg[1].pending ? 0 : 1 - t->stg_last);
  sb200_tracker::Staging& S = t->stg[k];
  // capacity exactly as predict() sizes it (hints included), so the predict call never reallocates a filled set
  const size_t T = (size_t)std::max<long long>(total, (long long)t->opts.max_scenes_hint * t->opts.max_dets_per_scene_hint);
  const size_t n = (size_t)total;
  int rc = 0;
  cudaStream_t cs = t->copy_stream;
  if (features == nullptr) has_feature = nullptr;
  // the set's previous reader (a frame that may still be in flight) finishes first; a reallocation meets the device
  if (S.boxes.bytes < T * 24 || (features && S.feat.bytes < T * (size_t)t->P.feature_dim * 4) || (has_feature && S.hasf.bytes < T) ||
      (quality && S.quality.bytes < T * 4) || (custom_ids && S.custom.bytes < T * 8) || (own_area && S.own.bytes < T * 4)) {
    if ((rc = t->drain())) return rc;
  }
  if (S.read_pending) { CU(cudaStreamWaitEvent(cs, S.ev_read, 0)); S.read_pending = false; }
  if ((rc = S.boxes.ensure(T * 24))) return rc;
  CU(cudaEventRecord(S.ev0, cs));
  CU(cudaMemcpyAsync(S.boxes.p, boxes, n * 24, cudaMemcpyHostToDevice, cs));
  if (features) {
    if ((rc = S.feat.ensure(T * (size_t)t->P.feature_dim * 4))) return rc;
    CU(cudaMemcpyAsync(S.feat.p, features, n * (size_t)t->P.feature_dim * 4, cudaMemcpyHostToDevice, cs));
    if (has_feature) { if ((rc = S.hasf.ensure(T))) return rc; CU(cudaMemcpyAsync(S.hasf.p, has_feature, n, cudaMemcpyHostToDevice, cs)); }
  }
  if (quality) { if ((rc = S.quality.ensure(T * 4))) return rc; CU(cudaMemcpyAsync(S.quality.p, quality, n * 4, cudaMemcpyHostToDevice, cs)); }
  if (custom_ids) { if ((rc = S.custom.ensure(T * 8))) return rc; CU(cudaMemcpyAsync(S.custom.p, custom_ids, n * 8, cudaMemcpyHostToDevice, cs)); }
  if (own_area) { if ((rc = S.own.ensure(T * 4))) return rc; CU(cudaMemcpyAsync(S.own.p, own_area, n * 4, cudaMemcpyHostToDevice, cs)); }
  CU(cudaEventRecord(S.ev, cs));
  S.key[0] = boxes; S.key[1] = features; S.key[2] = has_feature; S.key[3] = quality; S.key[4] = custom_ids; S.key[5] = own_area;
  S.total = total; S.pending = true;
  return 0;
}

int sb200_predict_batch_device(sb200_tracker* t, int32_t n_scenes, const uint64_t* scene_ids,
                               const int32_t* det_offsets, const float* boxes, const float* features,
                               const uint8_t* has_feature, const float* quality, const int64_t* custom_ids,
                               const float* own_area, const sb200_predict_out* out) {
  if (!t) return fail(SB200_ERR_INVALID, "tracker is NULL");
  return t->predict(n_scenes, scene_ids, det_offsets, boxes, features, has_feature, quality, custom_ids, own_area, out, true, false);
}

int sb200_skip_epochs(sb200_tracker* t, uint64_t scene_id, int32_t n) {
  if (!t || n < 0) return fail(SB200_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(t->device));
  { int rc_ = t->drain(); if (rc_) return rc_; }
  int slot = t->slot_for(scene_id, true);
  int rc = t->ensure_store((int)t->scene_of_slot.size(), std::max(t->track_cap, 64));
  if (rc) return rc;
  t->epoch[slot] += (uint32_t)n;
  return t->run_waste();  // skip_epochs_for_scene ends with auto_waste (tracker_api.rs:48-51)
}

int64_t sb200_current_epoch(sb200_tracker* t, uint64_t scene_id) {
  if (!t) return fail(SB200_ERR_INVALID, "tracker is NULL");
  int slot = t->slot_for(scene_id, false);
  return slot < 0 ? 0 : (int64_t)t->epoch[slot];
}

int64_t sb200_active_tracks(sb200_tracker* t) {
  if (!t) return fail(SB200_ERR_INVALID, "tracker is NULL");
  { int rc_ = t->drain(); if (rc_) return rc_; }
  int64_t n = 0;   // the reference's store still holds the expired tracks it has not collected yet
  for (int v : t->n_tracks) n += v;
  for (int v : t->n_hidden) n += v;
  return n;
}

int sb200_scene_track_counts(sb200_tracker* t, int32_t n_scenes, const uint64_t* scene_ids, int32_t* out) {
  if (!t || n_scenes < 0 || (n_scenes > 0 && (!scene_ids || !out))) return fail(SB200_ERR_INVALID, "bad arguments");
  { int rc_ = t->drain(); if (rc_) return rc_; }
  for (int s = 0; s < n_scenes; ++s) {
    int slot = t->slot_for(scene_ids[s], false);
    out[s] = slot < 0 ? 0 : t->n_tracks[slot] + t->n_hidden[slot];
  }
  return 0;
}

int sb200_scene_live_counts(sb200_tracker* t, int32_t n_scenes, const uint64_t* scene_ids, int32_t* live, int32_t* blocks) {
  if (!t || n_scenes < 0 || (n_scenes > 0 && !scene_ids)) return fail(SB200_ERR_INVALID, "bad arguments");
  { int rc_ = t->drain(); if (rc_) return rc_; }
  for (int s = 0; s < n_scenes; ++s) {
    int slot = t->slot_for(scene_ids[s], false);
    if (live) live[s] = slot < 0 ? 0 : t->n_tracks[slot];
    if (blocks) blocks[s] = slot < 0 ? 0 : (t->P.is_visual ? t->arena_top[slot] : t->n_tracks[slot]);
  }
  return 0;
}

int sb200_set_auto_waste(sb200_tracker* t, int32_t periodicity) {
  if (!t || periodicity < 0) return fail(SB200_ERR_INVALID, "bad arguments");
  t->auto_waste_periodicity = periodicity;
  t->auto_waste_counter = 0;  // set_auto_waste resets the counter (tracker_api.rs:29-33)
  return 0;
}

int sb200_clear_wasted(sb200_tracker* t) {
  if (!t) return fail(SB200_ERR_INVALID, "tracker is NULL");
  CU(cudaSetDevice(t->device));
  { int rc_ = t->drain(); if (rc_) return rc_; }
  // TrackerAPI::clear_wasted (src/trackers/tracker_api.rs:94-100) empties the wasted store; tracks swept early that the
  // reference has not collected yet are not in it and stay pending
  return t->drop_wasted_front(t->revealed);
}

int64_t sb200_wasted(sb200_tracker* t, int64_t cap, uint64_t* ids, uint64_t* scene_ids, uint32_t* epochs,
                     uint32_t* lengths, float* predicted_boxes, float* observed_boxes) {
  if (!t || cap < 0) return fail(SB200_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(t->device));
  { int rc_ = t->drain(); if (rc_) return rc_; }
  int rc = t->run_waste();  // wasted() starts with auto_waste (tracker_api.rs:90-91)
  if (rc) return rc;
  int64_t n = std::min<int64_t>(cap, t->wasted_count);
  if (n == 0) return 0;
  cudaStream_t st = t->stream;
  if (ids) CU(cudaMemcpyAsync(ids, t->wb.id, 8 * n, cudaMemcpyDeviceToHost, st));
  if (scene_ids) CU(cudaMemcpyAsync(scene_ids, t->wb.scene, 8 * n, cudaMemcpyDeviceToHost, st));
  if (epochs) CU(cudaMemcpyAsync(epochs, t->wb.epoch, 4 * n, cudaMemcpyDeviceToHost, st));
  if (lengths) CU(cudaMemcpyAsync(lengths, t->wb.length, 4 * n, cudaMemcpyDeviceToHost, st));
  if (predicted_boxes) CU(cudaMemcpyAsync(predicted_boxes, t->wb.pred, 24 * n, cudaMemcpyDeviceToHost, st));
  if (observed_boxes) CU(cudaMemcpyAsync(observed_boxes, t->wb.obs, 24 * n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if ((rc = t->drop_wasted_front(n))) return rc;   // drain
  return n;
}

int64_t sb200_wasted_history(sb200_tracker* t, int64_t cap, uint64_t* ids, uint64_t* scene_ids, uint32_t* epochs, uint32_t* lengths,
                             float* predicted_boxes, float* observed_boxes, int32_t history_cap, float* predicted_history,
                             float* observed_history, int32_t* history_counts) {
  if (!t || cap < 0 || history_cap < 0) return fail(SB200_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(t->device));
  { int rc_ = t->drain(); if (rc_) return rc_; }
  int rc = t->run_waste();  // wasted() starts with auto_waste (tracker_api.rs:90-91)
  if (rc) return rc;
  const int64_t n = std::min<int64_t>(cap, t->wasted_count);
  if (n == 0) return 0;
  cudaStream_t st = t->stream;
  std::vector<uint32_t> hlen((size_t)n);
  if (ids) CU(cudaMemcpyAsync(ids, t->wb.id, 8 * n, cudaMemcpyDeviceToHost, st));
  if (scene_ids) CU(cudaMemcpyAsync(scene_ids, t->wb.scene, 8 * n, cudaMemcpyDeviceToHost, st));
  if (epochs) CU(cudaMemcpyAsync(epochs, t->wb.epoch, 4 * n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hlen.data(), t->wb.length, 4 * n, cudaMemcpyDeviceToHost, st));
  std::vector<float> lastp((size_t)n * 6), lasto((size_t)n * 6);
  CU(cudaMemcpyAsync(lastp.data(), t->wb.pred, 24 * n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(lasto.data(), t->wb.obs, 24 * n, cudaMemcpyDeviceToHost, st));
  const int H = t->hist_len;
  std::vector<float> hp, ho;
  if (H > 1 && history_cap > 0) {
    hp.resize((size_t)n * H * 6); ho.resize((size_t)n * H * 6);
    CU(cudaMemcpyAsync(hp.data(), t->wb.hist_pred, sizeof(float) * hp.size(), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ho.data(), t->wb.hist_obs, sizeof(float) * ho.size(), cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  if (lengths) memcpy(lengths, hlen.data(), 4 * (size_t)n);
  if (predicted_boxes) memcpy(predicted_boxes, lastp.data(), 24 * (size_t)n);
  if (observed_boxes) memcpy(observed_boxes, lasto.data(), 24 * (size_t)n);
  // rings -> chronological order (oldest first), at most history_cap boxes per track
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t len = hlen[(size_t)i];
    int cnt = (int)std::min<uint32_t>(len, (uint32_t)H);
    cnt = std::min(cnt, (int)history_cap);
    if (history_counts) history_counts[i] = history_cap > 0 ? cnt : 0;
    for (int c = 0; c < cnt; ++c) {
      const uint32_t j = len - (uint32_t)cnt + (uint32_t)c;   // observation number
      const float* sp; const float* so;
      if (H > 1) { sp = &hp[((size_t)i * H + j % H) * 6]; so = &ho[((size_t)i * H + j % H) * 6]; }
      else { sp = &lastp[(size_t)i * 6]; so = &lasto[(size_t)i * 6]; }
      if (predicted_history) memcpy(predicted_history + ((size_t)i * history_cap + c) * 6, sp, 24);
      if (observed_history) memcpy(observed_history + ((size_t)i * history_cap + c) * 6, so, 24);
    }
  }
  if ((rc = t->drop_wasted_front(n))) return rc;   // drain
  return n;
}

static int64_t dump_scene(sb200_tracker* t, uint64_t scene_id, int64_t cap, bool idle_only, uint64_t* ids,
                          uint32_t* epochs, uint32_t* lengths, float* pred, float* obs, float* states30,
                          int32_t* feat_counts) {
  if (!t || cap < 0) return fail(SB200_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(t->device));
  { int rc_ = t->drain(); if (rc_) return rc_; }
  int slot = t->slot_for(scene_id, false);
  if (slot < 0) return 0;
  int n = t->n_tracks[slot];
  if (n == 0) return 0;
  size_t base = (size_t)slot * t->track_cap;
  std::vector<uint64_t> hid(n);
  std::vector<uint32_t> hep(n), hle(n);
  std::vector<float> hpr((size_t)n * 6), hob((size_t)n * 6), hst((size_t)n * 30);
  std::vector<unsigned char> hfc(n, 0);
  cudaStream_t st = t->stream;
  CU(cudaMemcpyAsync(hid.data(), t->ts.id + base, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hep.data(), t->ts.epoch + base, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hle.data(), t->ts.length + base, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hpr.data(), t->ts.pred + base * 6, 24 * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hob.data(), t->ts.obs + base * 6, 24 * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpy2DAsync(hst.data(), 120, t->ts.kst + base * sb::kStateStride, 4 * (size_t)sb::kStateStride, 120, (size_t)n, cudaMemcpyDeviceToHost, st));
  if (t->P.is_visual) CU(cudaMemcpyAsync(hfc.data(), t->ts.feat_cnt + base, (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  int64_t k = 0;
  for (int j = 0; j < n && k < cap; ++j) {
    // SortLookup::IdleLookup (src/trackers/sort.rs:190-208): last_updated_epoch != current epoch of the scene
    if (idle_only && hep[j] == t->epoch[slot]) continue;
    if (ids) ids[k] = hid[j];
    if (epochs) epochs[k] = hep[j];
    if (lengths) lengths[k] = hle[j];
    if (pred) memcpy(pred + k * 6, &hpr[(size_t)j * 6], 24);
    if (obs) memcpy(obs + k * 6, &hob[(size_t)j * 6], 24);
    if (states30) memcpy(states30 + k * 30, &hst[(size_t)j * 30], 120);
    if (feat_counts) feat_counts[k] = hfc[j];
    ++k;
  }
  return k;
}

// expired tracks of `scene_id` swept early: the reference's store still holds them (they are idle by definition)
static int64_t append_hidden(sb200_tracker* t, uint64_t scene_id, int64_t k, int64_t cap, uint64_t* ids, uint32_t* epochs,
                             uint32_t* lengths, float* pred, float* obs) {
  const int64_t h0 = t->revealed, hn = t->wasted_count - t->revealed;
  if (hn <= 0 || k >= cap) return k;
  std::vector<uint64_t> hid(hn), hsc(hn);
  std::vector<uint32_t> hep(hn), hle(hn);
  std::vector<float> hpr((size_t)hn * 6), hob((size_t)hn * 6);
  cudaStream_t st = t->stream;
  CU(cudaMemcpyAsync(hid.data(), t->wb.id + h0, 8 * (size_t)hn, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hsc.data(), t->wb.scene + h0, 8 * (size_t)hn, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hep.data(), t->wb.epoch + h0, 4 * (size_t)hn, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hle.data(), t->wb.length + h0, 4 * (size_t)hn, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hpr.data(), t->wb.pred + h0 * 6, 24 * (size_t)hn, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hob.data(), t->wb.obs + h0 * 6, 24 * (size_t)hn, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int64_t j = 0; j < hn && k < cap; ++j) {
    if (hsc[j] != scene_id) continue;
    if (ids) ids[k] = hid[j];
    if (epochs) epochs[k] = hep[j];
    if (lengths) lengths[k] = hle[j];
    if (pred) memcpy(pred + k * 6, &hpr[(size_t)j * 6], 24);
    if (obs) memcpy(obs + k * 6, &hob[(size_t)j * 6], 24);
    ++k;
  }
  return k;
}

int64_t sb200_idle_tracks(sb200_tracker* t, uint64_t scene_id, int64_t cap, uint64_t* ids, uint32_t* epochs,
                          uint32_t* lengths, float* predicted_boxes, float* observed_boxes) {
  int64_t k = dump_scene(t, scene_id, cap, true, ids, epochs, lengths, predicted_boxes, observed_boxes, nullptr, nullptr);
  if (k < 0) return k;
  return append_hidden(t, scene_id, k, cap, ids, epochs, lengths, predicted_boxes, observed_boxes);
}

int64_t sb200_scene_tracks(sb200_tracker* t, uint64_t scene_id, int64_t cap, uint64_t* ids, float* boxes,
                           float* states30, int32_t* feature_counts) {
  return dump_scene(t, scene_id, cap, false, ids, nullptr, nullptr, boxes, nullptr, states30, feature_counts);
}

int64_t sb200_last_costs(sb200_tracker* t, uint64_t scene_id, int64_t cap, float* out, int32_t* m, int32_t* n) {
  if (!t || !out || !m || !n) return fail(SB200_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(t->device));
  { int rc_ = t->drain(); if (rc_) return rc_; }
  *m = 0; *n = 0;
  if (t->last_n_scenes <= 0) return 0;
  // the scene table of the last frame was built on the device: read it back
  std::vector<sb::SceneDesc> sd((size_t)t->last_n_scenes);
  CU(cudaMemcpyAsync(sd.data(), t->f_scenes.p, sizeof(sb::SceneDesc) * sd.size(), cudaMemcpyDeviceToHost, t->stream));
  CU(cudaStreamSynchronize(t->stream));
  for (size_t si = 0; si < sd.size(); ++si) {
    const sb::SceneDesc& d = sd[si];
    if (d.scene_id != scene_id) continue;
    *m = d.m; *n = d.n;
    const int64_t full = (int64_t)d.m * d.n;
    const int64_t cnt = std::min<int64_t>(cap, full);
    if (cnt <= 0) return 0;
    // the dense matrix exists for scenes in dense voting mode and in SB200_FULL_COSTS runs; elsewhere the entry list is
    // the matrix (None wherever no entry is listed)
    int h[2] = {0, 0};   // pos_cnt, scene_mode
    const int ns = t->last_n_scenes;
    CU(cudaMemcpyAsync(&h[0], t->f_counters.as<int>() + si, 4, cudaMemcpyDeviceToHost, t->stream));
    CU(cudaMemcpyAsync(&h[1], t->f_counters.as<int>() + 2 * (size_t)ns + si, 4, cudaMemcpyDeviceToHost, t->stream));
    CU(cudaStreamSynchronize(t->stream));
    const bool dense_exists = h[1] != 0 || getenv("SB200_FULL_COSTS") != nullptr || getenv("SB200_NO_FORK") != nullptr;
    if (dense_exists) {
      CU(cudaMemcpyAsync(out, t->f_pos.as<float>() + d.pos_off, 4 * (size_t)cnt, cudaMemcpyDeviceToHost, t->stream));
      CU(cudaStreamSynchronize(t->stream));
      return cnt;
    }
    const int ne = std::min(h[0], d.pos_lcap);
    std::vector<sb::PosEntry> ents((size_t)std::max(ne, 0));
    if (ne > 0) {
      CU(cudaMemcpyAsync(ents.data(), t->f_poslist.as<sb::PosEntry>() + d.pos_lbase, sizeof(sb::PosEntry) * (size_t)ne,
                         cudaMemcpyDeviceToHost, t->stream));
      CU(cudaStreamSynchronize(t->stream));
    }
    const float qnan = std::nanf("");
    for (int64_t i = 0; i < cnt; ++i) out[i] = qnan;
    for (const sb::PosEntry& e : ents) {
      const int64_t idx = (int64_t)e.m * d.n + e.n;
      if (idx < cnt) out[idx] = e.v;
    }
    return cnt;
  }
  return 0;
}

int sb200_last_stage_ms(sb200_tracker* t, float* out5) {
  if (!t || !out5) return fail(SB200_ERR_INVALID, "bad arguments");
  { int rc_ = t->drain(); if (rc_) return rc_; }
  memcpy(out5, t->stage_ms, sizeof(float) * 5);
  return 0;
}

int sb200_last_kernel_ms(sb200_tracker* t, float* out2) {
  if (!t || !out2) return fail(SB200_ERR_INVALID, "bad arguments");
  { int rc_ = t->drain(); if (rc_) return rc_; }
  out2[0] = t->kernel_ms[0];
  out2[1] = t->kernel_ms[1];
  return 0;
}

void* sb200_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); g_err = "cudaMallocHost failed"; return nullptr; }
  return p;
}
void sb200_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
