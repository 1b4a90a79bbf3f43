// sb_engine.cuh -- device data layout and kernel launch interfaces of the association engine.
//
// HBM layout (all arrays allocated once per tracker and grown by doubling):
//   track store   : dense per scene slot, index = slot * track_cap + j, j in [0, n_tracks[slot]) in store order
//                   (insertion order; wasted tracks are removed by a stable compaction).  Array-of-structs rows
//                   of 6 / 30 floats so that a tile of tracks is one contiguous, fully coalesced block copy.
//   features      : [slot][track][physical obs slot][D8] f32, D8 = D rounded up to 8 lanes (Feature::from_vec
//                   zero-padding, src/track/utils.rs:45-71); the logical observation order of
//                   VisualMetric::optimize (src/trackers/visual_sort/metric.rs:297-374) is a per-track
//                   permutation (obs_phys) so feature rows never move.
//   per frame     : candidates in request order; cost matrices packed per scene (row-major m x n, and
//                   m x n x K for visual distances), NaN == None.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sb_math.cuh"

namespace sb {

constexpr int kMaxConstraints = 8;
constexpr int kMaxObs = 8;  // visual_max_observations supported on device (reference default 5)
constexpr int kStateStride = 32;   // floats per Kalman state row in the tracker's store (kStateFloats padded to 128 bytes)
constexpr int kMaxHist = 64;  // box history kept per track on the device (history_length above it, or 0 = unlimited, is capped)

struct Params {  // immutable per tracker, passed by value to kernels
  int kind, positional_kind, visual_kind;
  float iou_threshold, min_confidence, pos_weight, vel_weight;
  int max_idle_epochs;
  int n_constraints;
  int constraint_epochs[kMaxConstraints];
  float constraint_max_dist[kMaxConstraints];
  float visual_threshold;
  int feature_dim, d8, max_obs, min_votes, min_track_length;
  int vote_vis_cap;   // visual entries per scene the sparse voting kernel keeps in shared memory (0: kVoteVisCap)
  float min_area, min_quality_use, min_quality_collect, min_own_use, min_own_collect;
  bool is_visual, is_batch, use_own_area;
};

struct SceneDesc {  // one per scene of the current request
  int slot;        // scene slot in the track store
  int m;           // detections of this scene
  int n;           // stored tracks of this scene before this frame
  int det_base;    // first detection (row of the request)
  long long pos_off;  // offset of the m x n positional cost matrix
  long long vis_off;  // offset of the m x n x K visual matrix
  unsigned int epoch; // the scene's freshly incremented epoch (candidate epoch)
  int col_off;        // first entry of this scene in the per-frame column metadata (n * K physical feature rows)
  unsigned long long scene_id;
  int pos_lbase, pos_lcap;  // this scene's slice of the sparse positional entry list
  int vis_lbase, vis_lcap;  // this scene's slice of the visual survivor list
  int nb;             // feature blocks of this scene's arena in use (physical rows scanned by the screen = nb * K)
  int pad0;
  // dense tensor-core visual cost (kernels_feat_dense.cu) only:
  long long ws_off;   // first element of this scene's weight-sum matrix ws[block][candidate] (row pitch = m rounded up to 128)
  int blk_off;        // first entry of this scene in the per-block metadata (prefix of nb)
  int slab_off;       // first 256-column metadata slab of this scene (one slab per column tile)
};

// Host-written half of a scene descriptor.  The host knows the request (slot, detections, epoch, list slices) but -- with
// several frames in flight -- not how many tracks the scene's store holds when this frame runs; frame_setup_kernel joins
// the two on the device (n, nb, matrix / column / tile offsets) so that predict never waits for the previous frame.
struct SceneReq {
  int slot, m, det_base;
  unsigned int epoch;
  unsigned long long scene_id;
  int pos_lbase, pos_lcap, vis_lbase, vis_lcap;
};
struct FrameDyn {   // per-frame scalars only the device knows (written by frame_setup_kernel, read by later kernels)
  int n_tiles;            // tiles of the tensor-core visual cost kernel
  int total_cols;         // padded physical feature rows of all scenes (dense kernel: 256 x metadata slabs)
  int max_rows;           // max over the scenes of nb * K
  int max_n;              // max over the scenes of n
  long long pos_total;    // elements of the packed positional matrices
  long long vis_total;    // elements of the packed visual matrices
  unsigned long long units_mn;     // sum over the scenes of m * n   (pair-associations of this frame)
  unsigned long long units_rows;   // sum over the scenes of m * nb * K (dot products the visual cost kernel evaluates)
  long long live_total;   // sum over the scenes of n
  long long ws_total;     // elements of the dense kernel's weight-sum matrices
  int blk_total;          // arena blocks of all scenes
  int dense_scenes;       // scenes the dense exact kernels had to take (written late in the frame by scene_mode_kernel)
};

struct VisPair { int g, row, scene, outcol; };  // screen survivor: detection, feature row, scene, logical column
struct PosEntry { unsigned short m, n; float v; };  // one valid (candidate, track, cost) positional entry
constexpr int kVotePosCap = 3072;   // sparse entries per scene the voting kernel keeps in shared memory
constexpr int kVoteVisCap = 4096;   // power of two: the BestFit bitonic sort pads the list up to the next power of two

struct TrackStore {
  int track_cap;
  unsigned long long* id;
  unsigned int* epoch;
  unsigned int* length;
  long long* custom;
  signed char* vt;       // -1 == None
  float* pred;           // [idx][6] last predicted (posterior) box == observation attr box
  float* obs;            // [idx][6] last observed box
  float* radius;         // [idx]
  float* kst;            // [idx][kst_stride]: 30 state floats per track (8 mean + 8x8 covariance upper part ... see sb_math.cuh)
  int kst_stride;        // floats per row: 32 in the tracker's store (128-byte rows, 16-byte vector access), 30 for caller rows
  double* vert;          // [idx][8] vertex cache (IoU mode)
  // box history (SortAttributes::update_history, src/trackers/sort.rs:157-171): the last hist_len observed / predicted boxes
  // of every track as a ring, observation number j (0-based) in slot j % hist_len; null unless history_length > 1
  int hist_len;
  float* hist_pred;      // [idx][hist_len][6]
  float* hist_obs;       // [idx][hist_len][6]
  // visual
  float* feat;           // [idx][K][d8]
  void* feat_bf16;       // [idx][K][d8] bf16 copy of feat: B operand of the tensor-core screen
  float* fnorm2;         // [idx][K] by physical slot
  unsigned char* obs_phys;  // [idx][K] logical -> physical
  unsigned char* obs_hasf;  // [idx][K] logical: feature present
  float* obs_q;             // [idx][K] logical: quality
  unsigned char* obs_n;     // [idx]
  unsigned char* feat_cnt;  // [idx] visual_features_collected_count
  // Feature arena (tracker stores; null in the stateless operators, where block == track index).  A track owns one block
  // of K feature rows inside its scene's arena, feature row = (slot * track_cap + fblk[idx]) * K + physical slot.  The
  // small per-track arrays above are compacted (stably) whenever tracks expire; the feature rows never move: an expired
  // track's block goes to the scene's free list and is handed to the next new track.
  int* fblk;        // [idx] block of the track
  int* blk_owner;   // [slot * track_cap + block] track index j of the owner in current store order, -1: free
  int* blk_free;    // [slot * track_cap + i] free-list stack
  int* n_free;      // [slot]
  int* arena_top;   // [slot] blocks ever handed out (== live tracks + free blocks)
};

// first feature row of track `ti` (absolute store index) of scene slot `slot`, divided by K
__device__ __forceinline__ size_t feat_block(const TrackStore& ts, int slot, size_t ti) {
  return ts.fblk ? (size_t)slot * ts.track_cap + (size_t)ts.fblk[ti] : ti;
}

struct Frame {  // per-request transient device buffers (a request may be processed in scene chunks)
  int total;               // detections of this chunk
  int det0;                // first detection of this chunk (global row of the request)
  int scene0;              // first scene of this chunk (index into the request)
  const int* new_count_all;  // [all scenes of the request] (ids of non-batch trackers need the global prefix)
  long long pos_fill_off;  // offset of this chunk's positional matrices inside `pos`
  const float* in_boxes;   // [total][6] raw request boxes
  const float* in_feat;    // [total][D] or null
  const unsigned char* in_hasf;
  const float* in_quality;
  const long long* in_custom;
  const float* in_own;
  float* c_box;            // [total][6] candidate (Kalman-normalised) boxes
  float* c_radius;
  float* c_conf;           // max(conf, min_confidence)
  double* c_vert;          // [total][8]
  unsigned char* c_flags;  // bit0 has feature, bit1 feature usable (feature_can_be_used with *_use thresholds)
  float* c_norm2;
  void* c_bf16;            // [total][d8] bf16 copy of the candidate features (A operand of the screen)
  unsigned int* scene_max; // [n_scenes] order-preserving encoding of best.rs "max_dist"
  int* winner;             // [total] track index within the scene or -1
  unsigned char* c_vt;     // voting type of the decision
  float* pos;              // packed positional cost matrices
  // The dense positional matrices are only materialised where somebody reads them: the stateless operators and
  // SB200_FULL_COSTS runs (pos_dense_all), and the scenes that fall back to the dense voting kernel (filled and scanned
  // again after scene_mode is known).  Everywhere else the per-scene entry list IS the cost matrix.
  bool pos_dense_all;
  long long pos_total;     // elements in `pos` this frame
  float* vis;              // packed visual matrices
  SceneDesc* scenes;       // [n_scenes]
  int* new_count;          // [n_scenes] new tracks per scene (written by voting)
  int* status;             // [n_scenes] per-scene status flags (capacity overflow etc.)
  int* feat_dst;           // [total] destination feature row (block*K + phys) or -1
  int2* app_rank;          // [total] apply phase 1: (scene of the detection, rank among the scene's new tracks)
  int4* app_meta;          // [scenes] apply phase 1: (new tracks of earlier scenes, free blocks, arena top) before the frame
  int* frame_out;          // [n_scenes][3] written by the end-of-frame sweep: live tracks, arena blocks, newly expired
  const FrameDyn* dyn;     // device-built frame scalars (null in the stateless operators: host values are used)
  unsigned long long* id_counter;   // device copy of the tracker's id counter (null: the id_base argument is used)
  long long id_add;        // ids this frame consumes when known up front (batch trackers: one per detection), else -1
  // sparse views of the (mostly None) cost matrices, consumed by the voting stage
  PosEntry* pos_list;      // valid positional entries, per-scene slices
  int* pos_cnt;            // [n_scenes]
  VisPair* vis_pairs;      // screen survivors, per-scene slices
  float* vis_val;          // exact value of each survivor (NaN == failed the threshold)
  int* vis_cnt;            // [n_scenes]
  int* scene_mode;         // [n_scenes] 0: voting consumes the sparse lists; 1: dense matrices (a list overflowed)
  int* vis_mode;           // [n_scenes] visual side alone (set right after the screen): 0 = the survivors get refined
  // Lazy positional stage of the visual trackers (null: every pair is evaluated).  VisualVoting (visual_sort/voting.rs:
  // 45-100) only lets the positional metric decide candidates the visual BestFit pass left undecided, against tracks it
  // did not claim; a BestFit pre-pass publishes both sets and the culled scan skips everything else.
  unsigned char* decided;  // [total] candidate was decided by the visual pass
  unsigned char* excl;     // [slot * track_cap + n] track was claimed by the visual pass
  int* pre_winner;         // [total] the pre-pass's decision: track index the candidate won, -1 = decided as a new track
  int* dense_cnt;          // [1] scenes of the request in dense mode (null: unknown); lets the dense kernels leave at once
  int2* pos_gq;            // gated (candidate, track) pairs of the whole frame: x = scene, y = m << 16 | n (null: evaluate in the scan kernel)
  int* pos_gq_cnt;         // [1] entries of pos_gq (zeroed with the frame counters)
  int pos_gq_cap;
  int* refine_next;        // [n_scenes] next unclaimed survivor of the scene (the refinement's warps claim 32 at a time)
  const int* dense_bad;    // [n_scenes] dense tensor-core path only: != 0 sends the scene to the exact SIMT kernels
  // outputs (device), any may be null
  unsigned long long* o_ids;
  unsigned int* o_epochs;
  unsigned int* o_lengths;
  unsigned char* o_vt;
  float* o_pred;
  float* o_obs;
};

struct TcTile { int scene, m0, c0, pad; };  // pad: column-tile index inside the scene (dense kernel: its metadata slab)  // one 128 x 256 output tile of the tensor-core visual cost kernel
// per-frame metadata of one physical feature row (track n, physical slot p) of a scene, built once per frame
struct VisColMeta {
  float colb;    // column constant of the screen test (copy of TcArgs::colb)
  float colc;    // column part of the screen test
  int outcol;    // logical output column n*K + k (-1: none)
  int row;       // feature row idx*K + phys when the observation takes part in the metric, else -1
};
struct VisColGeo { float tx, ty, tr; unsigned int tep; };  // only read when spatio-temporal constraints exist
struct VisRowMeta { float rowk; int ok, pad0, pad1; };   // row constant of the screen test, candidate may vote visually
struct DenseTrackMeta { int n; int kt; float cmax; int pad; };   // arena block: owner (store index, -1: none), valid observations

// ---- kernel launchers (each in its own .cu) ----
void launch_prep(const Params& p, const Frame& f, int n_scenes, int max_m, cudaStream_t st);
// own-area shares of every detection among the detections of its scene (raw request boxes [total][6]) -> d_out[total].
// d_ovf_cnt / d_ovf ([total] (scene, detection) pairs): detections that more than kOwnMaxNb boxes overlap, handled by a
// second, CTA-per-detection pass; bit 1 of Frame::status[scene] is only set beyond kOwnBigNb overlapping boxes.
void launch_own_area(const Frame& f, int n_scenes, int max_m, const float* d_boxes, float* d_out, int* d_ovf_cnt,
                     int2* d_ovf, cudaStream_t st);
// dst (device) <- src (device alias of mapped pinned host memory), bytes a multiple of 4; a kernel instead of a DMA
void launch_pull(void* dst, const void* src, size_t bytes, cudaStream_t st);
// Per-frame tables built on the device: scene descriptors (request half from `req`, a device alias of mapped pinned host
// memory; store half from d_n_tracks / ts.arena_top), the tile list of the tensor-core visual cost kernel (mstep = 128 or
// 256 candidate rows per tile, 0: none) and the frame scalars; also zeroes the `n_zero` ints at `zero` (list counters, status).
// cstep: feature rows per column tile (256 for the screen; (256 / K) * K for the dense kernel, which also gets ws_off /
// blk_off / slab_off and one metadata slab per column tile instead of 128-padded columns)
void launch_frame_setup(const Params& p, const TrackStore& ts, const Frame& f, const SceneReq* req, int n_scenes,
                        const int* d_n_tracks, int mstep, int cstep, bool dense, TcTile* tiles, FrameDyn* dyn, int* zero,
                        int n_zero, cudaStream_t st);
// kernels launched by this library since it was loaded (every launch site counts itself)
void note_launch(int n = 1);
unsigned long long launch_count();
void launch_pos_cost(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_m, int max_n,
                     cudaStream_t st);
// visual cost: fp32 SIMT kernel in the reference's summation order (use_tc == false) or the tcgen05 3xTF32 kernel
struct TcArgs {
  int max_init_done;   // the frame's setup kernel already reset scene_max
  int colmeta_done;    // the column metadata of the screen was launched by the caller (side stream)  // tensor-core screen resources (all null / 0 when the dense exact kernel is used)
  bool use_tc;
  bool cluster2;   // tiles describe candidate-tile PAIRS processed by 2-CTA clusters (multicast B loads, or pair MMAs)
  bool pair;       // with cluster2: cta_group::2 MMAs (256 x 256 x 16 across the CTA pair)
  const TcTile* d_tiles;
  int n_tiles;            // tiles (stateless operators) or an upper bound of them (trackers: the count is d_n_tiles[0])
  const int* d_n_tiles;   // device-side tile count (null: n_tiles is exact)
  long long a_rows, b_rows;
  int num_sms;
  cudaEvent_t ev_screen0, ev_screen1, ev_refine1;  // optional per-kernel timing (null: not timed)
  VisColMeta* colmeta;   // [sum n_s*K]
  VisColGeo* colgeo;     // [sum n_s*K]
  float* colb;                // [sum n_s*K (+pad)] column constant of the screen test
  unsigned int* colvalid;     // bit per column: the observation takes part in the metric
  VisRowMeta* rowmeta;   // [total]
  int total_cols;
  int max_rows;          // max over the scenes of nb * K
  // ---- dense weight-sum path (mode 2)
  bool dense;            // run launch_vis_dense instead of the screen
  int cstep;             // feature rows per column tile: (256 / K) * K, so that no track straddles two tiles
  int max_blocks;        // upper bound of nb over the scenes
  int n_slabs_ub;        // upper bound of the 256-column metadata slabs of the frame
  void* ws;              // weight sums {S~, per-observation error bound} per (block, candidate), packed as half2
  unsigned int* d_rowb;  // [total][5] fused row bounds (per candidate and observation count)
  unsigned int* d_colb;  // [blk_ub] fused column bounds (per arena block of the frame)
  long long blk_ub;      // upper bound of the arena blocks of the frame
  float* slab_ktf;       // [slab][256] voting observations of the column's block (0: none)
  DenseTrackMeta* tmeta; // per arena block
  int2* rowinfo;         // per physical feature row: {logical output column, feature row or -1}
  float* slab_colc;      // [slab][256] column constant (|b|^2, or 1/|b| for cosine)
  float* slab_cmax;      // [slab][256] maximum of |b|^2 over the observations of the column's track
  unsigned int* slab_vmask;  // [slab][8] bit per column: observation takes part in the metric
  unsigned int* slab_bmask;  // [slab][8] bit per column: last physical slot of its block (flush the weight sum)
  float* scene_l0;       // [n_scenes] sampled lower bound of the scene's maximal distance (domain of the kernel's x)
  float* scene_cmax;     // [n_scenes] max |b|^2 of the scene
  VisPair* maxc;         // candidates for the maximal distance (same per-scene slices as the pair lists)
  float* maxc_val;
  int* maxc_cnt;         // [n_scenes]
  int* maxc_next;        // [n_scenes]
  int* dense_bad;        // [n_scenes] != 0: the dense result cannot be used for this scene (exact SIMT path takes it)
  int* zeros;            // [n_scenes] all zero (vis_mode view for the max refinement)
  int* dbg_counts;       // [8] per-frame diagnostics: scenes per fallback reason (1, 2, 4), max candidates
};
int launch_vis_cost(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_m, int max_n,
                    const TcArgs& tc, cudaStream_t st);
// Dense tensor-core visual cost for thresholds that cut nothing (the reference's default Euclidean(f32::MAX), the published
// bench's Euclidean(10.0) on unit vectors): see kernels_feat_dense.cu.  Fills the same per-scene pair lists the screen fills
// (whole (candidate, track) groups that can be a row or column maximum of BestFit's weight matrix), which the exact
// refinement and the sparse voting kernel then consume unchanged.
int launch_vis_dense(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_m, const TcArgs& tc,
                     cudaStream_t st);
// The same in two halves, so that the positional cost can run on a second stream next to the refinement:
//   _a: metadata, tensor-core screen, vis_mode, exact refinement of the survivors (needs nothing from the positional stage)
//   _b: final scene_mode (needs the positional list counters), dense exact kernel for the scenes in dense mode
int launch_vis_cost_a(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_m, int max_n,
                      const TcArgs& tc, cudaStream_t st);
int launch_vis_cost_b(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_m, int max_n,
                      const TcArgs& tc, cudaStream_t st);
// positional cost in two launches (dense None fill, culled scan) for callers that place them on a side stream
void launch_pos_fill(const Params& p, const Frame& f, int n_scenes, int max_m, int max_n, cudaStream_t st);
void launch_pos_scan(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_m, int max_n,
                     cudaStream_t st);
// pass 0: scenes whose visual lists are complete (vis_mode == 0) skip decided candidates / claimed tracks, the others are
// scanned in full; pass 1 (after scene_mode): full scan of the scenes that fell back to dense voting only then
void launch_pos_scan_lazy(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_m, int max_n,
                          int pass, cudaStream_t st);
// BestFit pre-pass of the sparse voting kernel: writes Frame::decided / Frame::excl for the scenes with vis_mode == 0
int launch_vote_masks(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_m, int max_n,
                      cudaStream_t st);
// phase 0: metadata + tensor-core screen; phase 1: exact refinement of the survivors of the sparse scenes
// screen metadata of the stored feature rows (needs the frame tables and the store, not the candidates)
void launch_vis_colmeta(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_n, const TcArgs& tc,
                        cudaStream_t st);
int launch_vis_cost_tc(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_n, const TcArgs& tc,
                       int phase, cudaStream_t st);
// materialises the dense visual matrix of the sparse scenes (operators / debugging only)
void launch_vis_densify(const Params& p, const Frame& f, int n_scenes, cudaStream_t st);
void launch_to_bf16(const float* src, int src_pitch, int d, int d8, long long rows, void* dst, cudaStream_t st);
// scene_max init (all scenes) and, unless init_only, the dense reduction for the scenes whose mode has bit1 set
void launch_scene_max(const Params& p, const Frame& f, int n_scenes, bool init_only, cudaStream_t st);
// per-scene voting mode from the list counters (runs after the cost kernels); launch_vis_mode: the visual half of it
void launch_scene_mode(const Params& p, const Frame& f, int n_scenes, bool tc_used, cudaStream_t st);
void launch_vis_mode(const Params& p, const Frame& f, int n_scenes, bool tc_used, cudaStream_t st);
// shared memory the voting kernels need for scenes of up to max_m x max_n (limit: kVotingSmemLimit)
size_t voting_smem_need(int max_m, int max_n, int viscap = 0);
constexpr size_t kVotingSmemLimit = 220 * 1024;
// returns cudaError from configuration (dynamic smem), 0 on success
int launch_voting(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_m, int max_n,
                  cudaStream_t st);
void launch_apply(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_m,
                  unsigned long long id_base, int* d_n_tracks, cudaStream_t st);
// the kept features of the frame -> the tracks' feature blocks; false when the frame has none (no launch)
bool launch_feat_store(const Params& p, const TrackStore& ts, const Frame& f, cudaStream_t st);
// stable compaction of wasted tracks; appends them to the wasted buffers
struct WastedBuf {
  int cap;
  int* count;  // device counter
  unsigned long long* id;
  unsigned long long* scene;
  unsigned int* epoch;
  unsigned int* length;
  float* pred;
  float* obs;
  float* hist_pred;   // [cap][hist_len][6] rings of the wasted tracks (null unless history_length > 1)
  float* hist_obs;
};
void launch_waste(const Params& p, const TrackStore& ts, int n_slots, const unsigned int* d_cur_epoch,
                  const unsigned long long* d_scene_ids, int* d_n_tracks, const WastedBuf& wb, int max_n,
                  cudaStream_t st);
// end-of-frame sweep over the scenes of the request: tracks that can never match again (EpochDb::baked,
// src/trackers/epoch_db.rs:51-66, evaluated at the scene's new epoch) leave the device store at once -- their records go
// to the wasted buffer, where the host keeps them hidden until the reference's own collection point -- and
// frame_out[s] = {live tracks, arena blocks, newly expired} is written for the host mirror.
void launch_frame_sweep(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int* d_n_tracks,
                        const WastedBuf& wb, cudaStream_t st);

// stateless operators
void launch_kalman_ops(int op, float pw, float vw, const float* in30, const float* boxes, int n, float* out30,
                       cudaStream_t st);
int launch_nms(const float* d_boxes, const float* d_scores, int n, float nms_thr, float score_thr, int has_score_thr,
               int* d_out_idx, int* d_out_count, cudaStream_t st);

// shared device helpers
__device__ __forceinline__ bool compat_ok(const Params& p, unsigned int cand_epoch, unsigned int trk_epoch, float cx,
                                          float cy, float cr, float tx, float ty, float tr) {
  // SortAttributes::compatible, src/trackers/sort.rs:250-270 (scene equality is structural here)
  unsigned int delta = cand_epoch > trk_epoch ? cand_epoch - trk_epoch : trk_epoch - cand_epoch;
  if ((unsigned int)p.max_idle_epochs < delta) return false;
  // SpatioTemporalConstraints::validate, src/trackers/spatio_temporal_constraints.rs:48-59
  for (int i = 0; i < p.n_constraints; ++i) {
    if ((unsigned int)p.constraint_epochs[i] >= delta) {
      float d = dist_in_2r(cx, cy, cr, tx, ty, tr);
      return d <= p.constraint_max_dist[i];
    }
  }
  return true;
}

}  // namespace sb
