/*
 * similari_b200.h -- C ABI of the B200-native association engine (libsimilari_b200.so).
 *
 * Drop-in boundary for Similari's per-frame cost-matrix + assignment hot path.  The reference has no C ABI
 * (it is a Rust crate with PyO3 classes); these entry points are what a Rust `extern "C"` block inside
 * Sort / BatchSort / VisualSort / BatchVisualSort (or the ctypes/PyO3 layer) binds in place of
 *     TrackStore::foreign_track_distances  (src/track/store.rs:429-460)
 *   + Voting::winners                       (src/trackers/sort/voting.rs:30-100, src/trackers/visual_sort/voting.rs:45-100)
 *   + TrackStore::merge_external / add_track (src/track/store.rs:625-691)
 * INTEGRATION.md shows the Rust-side and Python-side bindings.  All paths cited below are relative to the
 * reference tree (insight-platform/Similari, crate similari-trackers-rs v0.26.12).
 *
 * Conventions
 *   - plain pointers and sizes only; the caller allocates every output; the library never frees caller memory;
 *   - a box is 6 floats (xc, yc, angle, aspect, height, confidence) = Universal2DBox (src/utils/bbox.rs:79-87);
 *     angle == NaN encodes Option::None;
 *   - custom_object_id == INT64_MIN encodes Option::None; feature quality NULL means 1.0 (unwrap_or(1.0),
 *     src/trackers/visual_sort/simple_api.rs:141-151);
 *   - every function returns 0 on success, a negative sb200_status otherwise (the reference panics instead);
 *     sb200_last_error() returns the message of the calling thread's last failure;
 *   - a tracker handle is single-threaded (`&mut self` in the reference); handles are independent;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with SB200_ERR_CUDA.
 */
#ifndef SIMILARI_B200_H
#define SIMILARI_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SB200_VERSION 2

typedef enum {
  SB200_OK = 0,
  SB200_ERR_INVALID = -1,   /* bad argument (the reference's assert!/panic paths) */
  SB200_ERR_CUDA = -2,      /* CUDA runtime error or no device */
  SB200_ERR_CAPACITY = -3,  /* problem exceeds a documented device limit */
  SB200_ERR_INTERNAL = -4
} sb200_status;

/* TrackerKind: which reference tracker's semantics the handle follows (id numbering, voting cascade).
 *   SORT             src/trackers/sort/simple_api.rs:21-29
 *   BATCH_SORT       src/trackers/sort/batch_api.rs:46-53
 *   VISUAL_SORT      src/trackers/visual_sort/simple_api.rs
 *   BATCH_VISUAL_SORT src/trackers/visual_sort/batch_api.rs */
#define SB200_KIND_SORT 0
#define SB200_KIND_BATCH_SORT 1
#define SB200_KIND_VISUAL_SORT 2
#define SB200_KIND_BATCH_VISUAL_SORT 3
/* PositionalMetricType, src/trackers/sort.rs:364-369 */
#define SB200_POS_MAHA 0
#define SB200_POS_IOU 1
/* VisualSortMetricType, src/trackers/visual_sort/metric.rs:20-24 */
#define SB200_VIS_EUCLIDEAN 0
#define SB200_VIS_COSINE 1
/* VotingType, src/trackers/sort.rs:357-362 */
#define SB200_VOTING_VISUAL 0
#define SB200_VOTING_POSITIONAL 1
#define SB200_MAX_CONSTRAINTS 8
#define SB200_NONE_ID INT64_MIN

/* Constructor arguments of the four trackers folded into one struct:
 *   Sort::new / BatchSort::new              src/trackers/sort/simple_api.rs:41-50, batch_api.rs:157-167
 *   VisualSortOptions + VisualMetricBuilder src/trackers/visual_sort/options.rs:10-205, metric/builder.rs:9-42 */
typedef struct {
  int32_t kind;
  int32_t positional_kind;        /* method / positional_metric */
  float iou_threshold;            /* PositionalMetricType::IoU(t) */
  float min_confidence;           /* min_confidence / positional_min_confidence */
  int32_t max_idle_epochs;
  int32_t history_length;         /* bbox_history / kept_history_length: boxes of history per track (device cap 64; 0 = that cap) */
  float kalman_position_weight;
  float kalman_velocity_weight;
  int32_t n_constraints;          /* SpatioTemporalConstraints: (epoch_delta, max_distance) pairs */
  int32_t constraint_epochs[SB200_MAX_CONSTRAINTS];
  float constraint_max_dist[SB200_MAX_CONSTRAINTS];
  int32_t visual_kind;
  float visual_threshold;
  int32_t feature_dim;            /* D; 0 for Sort / BatchSort */
  int32_t visual_max_observations;
  int32_t visual_min_votes;
  int32_t visual_minimal_track_length;
  float visual_minimal_area;
  float visual_minimal_quality_use;
  float visual_minimal_quality_collect;
  float visual_minimal_own_area_percentage_use;
  float visual_minimal_own_area_percentage_collect;
  int32_t max_scenes_hint;            /* capacity hints (0 = grow on demand) */
  int32_t max_tracks_per_scene_hint;
  int32_t max_dets_per_scene_hint;
  int32_t device;                     /* CUDA device ordinal */
} sb200_options;

/* Fills `o` with the reference's defaults (PySort::new defaults, src/trackers/sort/simple_api.rs:461-470;
 * VisualMetricBuilder::default, src/trackers/visual_sort/metric/builder.rs:26-42). */
void sb200_options_default(sb200_options* o);

typedef struct sb200_tracker sb200_tracker;

/* Per-detection result columns = SortTrack (src/trackers/sort.rs:286-311) as struct-of-arrays.  Any pointer
 * may be NULL (that column is then neither computed into host memory nor copied back). */
typedef struct {
  uint64_t* ids;            /* SortTrack.id */
  uint32_t* epochs;         /* SortTrack.epoch */
  uint32_t* lengths;        /* SortTrack.length */
  uint8_t* voting_types;    /* SortTrack.voting_type (SB200_VOTING_*) */
  float* predicted_boxes;   /* [total][6] SortTrack.predicted_bbox */
  float* observed_boxes;    /* [total][6] SortTrack.observed_bbox */
} sb200_predict_out;

const char* sb200_last_error(void);
int sb200_device_count(void);

/* ---- tracker lifecycle (Sort::new ... ) ---- */
int sb200_tracker_create(const sb200_options* opts, sb200_tracker** out);
void sb200_tracker_destroy(sb200_tracker* t);
/* Orders the tracker's work with `cuda_stream` (a cudaStream_t; NULL is the legacy default stream): every predict call first waits
 * for what that stream holds when the call is made (so device-resident inputs may be produced on it just before the call)
 * and the stream waits for the call's frame (so its outputs can be consumed on it, and the caller can bracket the work
 * with its own CUDA events).  The kernels themselves run on the tracker's own streams. */
int sb200_tracker_set_stream(sb200_tracker* t, void* cuda_stream);
/* per_call != 0 (the default): the caller's stream waits for the frame of every predict call.  0: it does not -- successive
 * frames then overlap where they can (the next frame's candidate preparation runs under the current frame's cost kernels);
 * a stream that consumes device-resident outputs -- or overwrites device-resident inputs of a frame that may still be
 * running -- waits explicitly with sb200_stream_join (or the host calls sb200_sync).  Counterpart of taking the results off PredictionBatchResult's channel when they are needed
 * (src/trackers/batch.rs:24-38) instead of blocking in predict. */
int sb200_set_stream_join(sb200_tracker* t, int32_t per_call);
/* Makes `cuda_stream` wait (on the device) for every frame enqueued so far. */
int sb200_stream_join(sb200_tracker* t, void* cuda_stream);

/* VisualSortObservation::feature is an Option (src/trackers/visual_sort.rs:42-55): a tracker may see frames without any
 * feature before it learns the feature length.  Until a request has carried feature rows the dimension given at creation
 * is provisional and can be changed here; afterwards a different value is SB200_ERR_INVALID. */
int sb200_set_feature_dim(sb200_tracker* t, int32_t feature_dim);

/* ---- the hot path ----
 * One call == Sort::predict_with_scene (n_scenes = 1, src/trackers/sort/simple_api.rs:110-196) or
 * BatchSort::predict / BatchVisualSort::predict over a PredictionBatchRequest (src/trackers/sort/batch_api.rs:222-290,
 * src/trackers/visual_sort/batch_api.rs:213-317, src/trackers/batch.rs:12-38) flattened as:
 *   scene_ids[n_scenes], det_offsets[n_scenes+1] (CSR over detections),
 *   boxes[total][6], features[total][D] or NULL, has_feature[total] or NULL (all present),
 *   quality[total] or NULL, custom_ids[total] or NULL, own_area[total] or NULL
 *   (own-area shares of exclusively_owned_areas, computed by the caller; src/utils/clipping/bbox_own_areas.rs).
 * All pointers are HOST pointers; inputs are copied to the device and the requested result columns copied back
 * before the call returns (results in input order, like the reference's Vec<SortTrack>). */
int sb200_predict_batch(sb200_tracker* t, int32_t n_scenes, const uint64_t* scene_ids, const int32_t* det_offsets,
                        const float* boxes, const float* features, const uint8_t* has_feature, const float* quality,
                        const int64_t* custom_ids, const float* own_area, const sb200_predict_out* out);
/* Optional input prefetch for pipelined callers (BatchSort::predict is asynchronous in the reference too,
 * src/trackers/sort/batch_api.rs:222-290): starts the host-to-device copy of a FUTURE request's columns on a copy
 * stream and returns immediately.  A later sb200_predict_batch() called with the same `boxes` / `features` pointers
 * and the same detection count uses the prefetched copy instead of copying again, so the copy of frame i+1 overlaps
 * the kernels of frame i.  The host buffers must stay unchanged until that predict call returns. */
int sb200_prefetch_inputs(sb200_tracker* t, int32_t total, const float* boxes, const float* features,
                          const uint8_t* has_feature, const float* quality, const int64_t* custom_ids,
                          const float* own_area);
/* The asynchronous form of the same call -- what BatchSort::predict is in the reference, where the request is queued
 * and the per-scene results arrive later on PredictionBatchResult's channel (src/trackers/sort/batch_api.rs:222-290,
 * src/trackers/batch.rs:24-38).  The frame is enqueued on the tracker's stream and the call returns; up to three frames
 * are in flight (a fourth call waits for the oldest).  The host buffers (inputs, and the `out` columns, which should be
 * pinned: sb200_host_alloc) must stay valid and are only defined after sb200_sync() -- or after a later call has
 * reported sb200_frames_in_flight() low enough.  An error inside an asynchronous frame is returned by the next
 * predict / sync / query call on the tracker. */
int sb200_predict_batch_async(sb200_tracker* t, int32_t n_scenes, const uint64_t* scene_ids, const int32_t* det_offsets,
                              const float* boxes, const float* features, const uint8_t* has_feature, const float* quality,
                              const int64_t* custom_ids, const float* own_area, const sb200_predict_out* out);
/* Waits for every frame in flight (PredictionBatchResult::get until batch_size results arrived) and hands their
 * bookkeeping to the host side of the tracker; returns the first error an asynchronous frame raised, if any. */
int sb200_sync(sb200_tracker* t);
/* Frames enqueued and not yet completed (PredictionBatchResult::ready is `== 0`); never blocks. */
int sb200_frames_in_flight(sb200_tracker* t);

/* Host-side cost of the predict entry points since the tracker was created: out3 = {calls, wall milliseconds spent inside
 * them, milliseconds of that spent blocked on the device (ring of frames in flight full, wait == true, reallocation)}.
 * (total - blocked) / calls is what one frame costs the calling thread; the stream-ordered design needs it below the
 * frame's device time (no reference counterpart: the reference's predict is synchronous). */
int sb200_host_counters(sb200_tracker* t, double* out3);
/* Same call with boxes / features / has_feature / quality / custom_ids / own_area and every non-NULL `out` column
 * being DEVICE pointers (inputs already resident in HBM).  scene_ids and det_offsets stay host pointers (they are
 * consumed before the call returns).  Stream-ordered like sb200_predict_batch_async: nothing in the call waits for
 * the device -- the per-frame tables that depend on the previous frame's outcome (tracks per scene, matrix offsets, the
 * tile list of the tensor-core kernel, the id counter) are built by a kernel -- so consecutive frames queue back to back
 * and a caller can order its own work after the results with an event on the tracker's stream. */
int sb200_predict_batch_device(sb200_tracker* t, int32_t n_scenes, const uint64_t* scene_ids,
                               const int32_t* det_offsets, const float* boxes, const float* features,
                               const uint8_t* has_feature, const float* quality, const int64_t* custom_ids,
                               const float* own_area, const sb200_predict_out* out);

/* ---- TrackerAPI (src/trackers/tracker_api.rs:27-117) ---- */
int sb200_skip_epochs(sb200_tracker* t, uint64_t scene_id, int32_t n);
int64_t sb200_current_epoch(sb200_tracker* t, uint64_t scene_id);
int64_t sb200_active_tracks(sb200_tracker* t);                 /* sum(active_shard_stats()) */
/* stored tracks per scene (0 for unknown scenes) as the reference's store would count them: the N of the next frame's
 * N x M cost matrix, expired tracks that the reference has not collected yet included */
int sb200_scene_track_counts(sb200_tracker* t, int32_t n_scenes, const uint64_t* scene_ids, int32_t* out);
/* what the device actually holds and scans per scene: `live` = tracks that can still match (expired tracks leave the
 * device store at the end of the frame in which they expire and wait, hidden, for the reference's collection point:
 * auto-waste tick, wasted(), skip_epochs), `blocks` = feature blocks of the scene's arena (rows scanned by the visual
 * cost kernel = blocks * visual_max_observations).  Either output may be NULL. */
int sb200_scene_live_counts(sb200_tracker* t, int32_t n_scenes, const uint64_t* scene_ids, int32_t* live, int32_t* blocks);
int sb200_set_auto_waste(sb200_tracker* t, int32_t periodicity);
int sb200_clear_wasted(sb200_tracker* t);
/* wasted(): drains up to `cap` wasted tracks; returns the count (>= 0) or a negative status. */
int64_t sb200_wasted(sb200_tracker* t, int64_t cap, uint64_t* ids, uint64_t* scene_ids, uint32_t* epochs,
                     uint32_t* lengths, float* predicted_boxes, float* observed_boxes);
/* wasted() with the box history of WastedSortTrack (src/trackers/sort.rs:316-341: predicted_boxes / observed_boxes, the last
 * history_length boxes of the track, oldest first, kept by SortAttributes::update_history, sort.rs:157-171).
 * predicted_history / observed_history: [cap][history_cap][6], history_counts[cap] = boxes filled for each track. */
int64_t sb200_wasted_history(sb200_tracker* t, int64_t cap, uint64_t* ids, uint64_t* scene_ids, uint32_t* epochs,
                             uint32_t* lengths, float* predicted_boxes, float* observed_boxes, int32_t history_cap,
                             float* predicted_history, float* observed_history, int32_t* history_counts);
/* idle_tracks_with_scene() (src/trackers/sort/simple_api.rs:198-215) */
int64_t sb200_idle_tracks(sb200_tracker* t, uint64_t scene_id, int64_t cap, uint64_t* ids, uint32_t* epochs,
                          uint32_t* lengths, float* predicted_boxes, float* observed_boxes);
/* Debug / parity: dense dump of one scene's store in store order.  states: [n][30] = mean[10] + 5 x (Pxx,Pxv,Pvx,Pvv). */
int64_t sb200_scene_tracks(sb200_tracker* t, uint64_t scene_id, int64_t cap, uint64_t* ids, float* boxes,
                           float* states30, int32_t* feature_counts);
/* Debug / parity: last frame's positional cost matrix of a scene ([m][n] f32, NaN == None), n = the scene's live tracks
 * in store order.  Visual trackers on the tensor-core path evaluate the positional metric lazily -- only for candidates
 * the visual BestFit pass left undecided against tracks it did not claim, the pairs VisualVoting::winners
 * (src/trackers/visual_sort/voting.rs:45-100) can still consult -- so the other entries read None; with the environment
 * variable SB200_FULL_COSTS=1 every pair is evaluated as the reference does. */
int64_t sb200_last_costs(sb200_tracker* t, uint64_t scene_id, int64_t cap, float* out, int32_t* m, int32_t* n);
/* Cumulative work of all completed frames (waits for the frames in flight): counters4 = { sum over frames and scenes of
 * M x N (pair-associations, N = tracks the device store held when the frame ran), sum of M x (feature rows scanned by the
 * visual cost kernel), frames, scenes the exact SIMT fallback kernels had to take }, ms8 = { summed per-stage device times: prep, positional cost, visual cost, voting, apply;
 * summed times of the dominant visual-cost kernel and of the refinement; frames in which those two ran }.  Either may
 * be NULL.  bench.py reads it before and after its timed region. */
int sb200_work_counters(sb200_tracker* t, uint64_t* counters4, double* ms8);
/* Kernels this library has launched since it was loaded (every launch site counts itself). */
uint64_t sb200_launch_count(void);
/* Per-stage device times (ms) of the last completed predict call: prep, positional cost, visual cost, voting, apply. */
int sb200_last_stage_ms(sb200_tracker* t, float* out5);
/* Device times (ms) of the dominant visual-cost kernels of the last predict call: [0] tensor-core screen kernel,
 * [1] scene-mode + exact refinement kernels; 0 when the tensor-core path was not used. */
int sb200_last_kernel_ms(sb200_tracker* t, float* out2);

/* ---- stateless operators (host pointers) used by parity tests and by callers that keep their own state ----
 * Positional cost matrix = SortMetric::metric over all pairs (src/trackers/sort/metric.rs:38-77):
 * out[m][n] = IoU*conf (>= thr) or (100 - d^2)/conf, NaN == None.  track_states30 only for Mahalanobis. */
int sb200_sort_cost_matrix(int32_t positional_kind, float iou_threshold, float min_confidence, float pos_weight,
                           float vel_weight, const float* cand_boxes, int32_t m, const float* track_boxes,
                           const float* track_states30, int32_t n, float* out_mn, int32_t device);
/* Visual cost matrix = euclidean / cosine (src/distance.rs:9-47) + is_ok/distance_to_weight
 * (src/trackers/visual_sort/metric.rs:52-64); out[m][n], NaN == None. */
int sb200_visual_cost_matrix(int32_t visual_kind, float threshold, const float* cand_features, int32_t m,
                             const float* track_features, int32_t n, int32_t d, float* out_mn, int32_t device);
/* SortVoting::winners on a dense cost matrix (NaN == None): winner[m] = track index or -1 (new track). */
int sb200_sort_voting(float threshold, const float* cost_mn, int32_t m, int32_t n, int32_t* winner, int32_t device);
/* VisualVoting::winners on dense matrices: pos[m][n], vis[m][n][k] (NaN == None). */
int sb200_visual_voting(float positional_threshold, int32_t min_votes, const float* pos_mn, const float* vis_mnk,
                        int32_t m, int32_t n, int32_t k, int32_t* winner, uint8_t* voting_type, int32_t device);
/* Kalman filter steps on packed states (src/utils/kalman/kalman_2d_box.rs:58-148), n states at once. */
int sb200_kalman_initiate(float pos_weight, float vel_weight, const float* boxes, int32_t n, float* states30, int32_t device);
int sb200_kalman_predict(float pos_weight, float vel_weight, const float* in30, int32_t n, float* out30, int32_t device);
int sb200_kalman_update(float pos_weight, float vel_weight, const float* in30, const float* boxes, int32_t n,
                        float* out30, int32_t device);
/* exclusively_owned_areas + exclusively_owned_areas_normalized_shares (src/utils/clipping/bbox_own_areas.rs:8-46) for the
 * boxes of ONE scene: out[i] = share of box i that no other box covers, in [0, 1].  The visual trackers call the same
 * kernel themselves when an own-area threshold is set and the request carries no `own_area` column
 * (src/trackers/visual_sort/simple_api.rs:110-127).  A box that more than 32 others overlap takes a second,
 * CTA-per-box pass; SB200_ERR_CAPACITY only beyond 2800 overlapping boxes on one box. */
int sb200_own_area_shares(const float* boxes, int32_t n, float* out, int32_t device);

/* nms (src/utils/nms.rs:32-72): scores NULL or NaN entries == None; out_idx = kept input indices in rank order;
 * returns kept count or negative status. */
int64_t sb200_nms(const float* boxes, const float* scores, int32_t n, float nms_threshold, float score_threshold,
                  int32_t has_score_threshold, int32_t* out_idx, int32_t device);

/* ---- multi-GPU: the one exchange step of the scene-sharded path (csrc/comm.cu) ----
 * Scenes are independent and track state is sticky per GPU (rank = scene shard), so N GPUs run N independent trackers; the
 * only data that crosses GPUs is the request on its way from an ingest rank to the owners of its scenes and the assigned
 * track records on their way back -- the counterpart of the reference's voting-shard fan-out and result channel
 * (src/trackers/sort/batch_api.rs:197-207,222-290).  One process per GPU.  NCCL (send/recv over NVLink) is loaded at run
 * time; rank 0 creates the 128-byte unique id and the caller ships it to the other ranks on its own control channel.
 * All data pointers are DEVICE pointers; calls are asynchronous on `cuda_stream`, so the scatter of frame i+1 can overlap
 * the kernels of frame i on another stream.  det_range[world + 1]: rank r owns detections [det_range[r], det_range[r+1])
 * of the root's request (its scenes' detections are contiguous).  Non-root ranks pass NULL for the `all_*` arguments. */
typedef struct sb200_comm sb200_comm;
int sb200_comm_unique_id(void* out128);
int sb200_comm_create(int32_t rank, int32_t world, const void* id128, int32_t device, sb200_comm** out);
void sb200_comm_destroy(sb200_comm* c);
/* root -> owners: boxes [6 f32], features [feature_dim f32], has_feature, quality, custom ids; a column is skipped on every
 * rank when its `my_*` pointer is NULL (all ranks must agree). */
int sb200_shard_scatter(sb200_comm* c, int32_t root, const int32_t* det_range, int32_t feature_dim, const float* all_boxes,
                        const float* all_features, const uint8_t* all_has_feature, const float* all_quality,
                        const int64_t* all_custom_ids, float* my_boxes, float* my_features, uint8_t* my_has_feature,
                        float* my_quality, int64_t* my_custom_ids, void* cuda_stream);
/* owners -> root: the SortTrack columns (`mine`: this rank's results as written by sb200_predict_batch_device; `all`: the
 * root's buffers for the whole request; a column is skipped when `mine` has it NULL). */
int sb200_shard_gather(sb200_comm* c, int32_t root, const int32_t* det_range, const sb200_predict_out* mine,
                       const sb200_predict_out* all, void* cuda_stream);

/* Pinned host memory for callers that want the predict H2D/D2H copies to run at full PCIe speed. */
void* sb200_host_alloc(size_t bytes);
void sb200_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
