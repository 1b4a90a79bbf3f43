"""Array-level Python interface of the engine: numpy in, numpy out, one C-ABI call per frame.

This is the zero-glue path the benchmark and the parity tests use; `similari_b200.api` layers the reference's
PyO3 class names (Sort, BatchSort, VisualSort, BatchVisualSort, ...) on top of it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Options, PredictOut, check, default_options, lib, ptr

F32MAX = float(np.finfo(np.float32).max)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Tracker:
    """Device-resident Sort / BatchSort / VisualSort / BatchVisualSort (selected by opts.kind)."""

    def __init__(self, opts: Options):
        self.opts = opts
        self._L = lib()
        h = C.c_void_p()
        check(self._L.sb200_tracker_create(C.byref(opts), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.sb200_tracker_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int, join_per_call: bool = True):
        """Orders every call after what `cuda_stream` holds at that moment; with join_per_call the stream also waits for
        each call's frame (else use stream_join / sync before consuming device-resident outputs)."""
        check(self._L.sb200_tracker_set_stream(self._h, C.c_void_p(cuda_stream)))
        check(self._L.sb200_set_stream_join(self._h, 1 if join_per_call else 0))

    def stream_join(self, cuda_stream: int):
        """sb200_stream_join: `cuda_stream` waits on the device for every frame enqueued so far."""
        check(self._L.sb200_stream_join(self._h, C.c_void_p(cuda_stream)))

    def predict_batch(self, scene_ids, det_offsets, boxes, features=None, has_feature=None, quality=None,
                      custom_ids=None, own_area=None, want=("ids", "epochs", "lengths", "voting_types", "predicted",
                                                            "observed"), out=None, wait=True):
        """Host-pointer call (sb200_predict_batch).  Returns a dict of numpy arrays (the SortTrack columns).
        wait=False: sb200_predict_batch_async -- the arrays (pass pinned ones in `out`) are defined after sync()."""
        scene_ids = np.ascontiguousarray(scene_ids, dtype=np.uint64)
        det_offsets = np.ascontiguousarray(det_offsets, dtype=np.int32)
        total = int(det_offsets[-1]) if len(det_offsets) else 0
        boxes = _f32(boxes).reshape(-1, 6)
        if len(boxes) != total:
            raise ValueError("boxes rows != det_offsets[-1]")
        features = _f32(features) if features is not None else None
        has_feature = np.ascontiguousarray(has_feature, dtype=np.uint8) if has_feature is not None else None
        quality = _f32(quality) if quality is not None else None
        custom_ids = np.ascontiguousarray(custom_ids, dtype=np.int64) if custom_ids is not None else None
        own_area = _f32(own_area) if own_area is not None else None
        if out is None:
            out = {}
            if "ids" in want:
                out["ids"] = np.zeros(total, dtype=np.uint64)
            if "epochs" in want:
                out["epochs"] = np.zeros(total, dtype=np.uint32)
            if "lengths" in want:
                out["lengths"] = np.zeros(total, dtype=np.uint32)
            if "voting_types" in want:
                out["voting_types"] = np.zeros(total, dtype=np.uint8)
            if "predicted" in want:
                out["predicted"] = np.zeros((total, 6), dtype=np.float32)
            if "observed" in want:
                out["observed"] = np.zeros((total, 6), dtype=np.float32)
        po = PredictOut(ptr(out.get("ids")), ptr(out.get("epochs")), ptr(out.get("lengths")),
                        ptr(out.get("voting_types")), ptr(out.get("predicted")), ptr(out.get("observed")))
        fn = self._L.sb200_predict_batch if wait else self._L.sb200_predict_batch_async
        check(fn(self._h, len(scene_ids), ptr(scene_ids), ptr(det_offsets), ptr(boxes), ptr(features), ptr(has_feature),
                 ptr(quality), ptr(custom_ids), ptr(own_area), C.byref(po)))
        if not wait:   # the caller's arrays must outlive the frame
            self._keep = getattr(self, "_keep", [])[-16:] + [(boxes, features, has_feature, quality, custom_ids, own_area, out)]
        return out

    def set_feature_dim(self, dim):
        """sb200_set_feature_dim: fixes the feature length of a visual tracker that has not stored a feature yet."""
        check(self._L.sb200_set_feature_dim(self._h, int(dim)))

    def sync(self):
        """sb200_sync: waits for every frame in flight; raises the first error an asynchronous frame produced."""
        check(self._L.sb200_sync(self._h))

    def frames_in_flight(self):
        return int(check(self._L.sb200_frames_in_flight(self._h)))

    def work_counters(self):
        """Cumulative work of the completed frames (waits for the frames in flight)."""
        c = np.zeros(4, np.uint64)
        ms = np.zeros(8, np.float64)
        check(self._L.sb200_work_counters(self._h, ptr(c), ptr(ms)))
        return {"pair_associations": int(c[0]), "visual_dot_products": int(c[1]), "frames": int(c[2]), "dense_fallback_scenes": int(c[3]),
                "stage_ms": dict(zip(("prep", "positional_cost", "visual_cost", "voting", "apply"), map(float, ms[:5]))),
                "vis_screen_ms": float(ms[5]), "vis_refine_ms": float(ms[6]), "tc_frames": int(ms[7])}

    def host_counters(self):
        """sb200_host_counters: calls of the predict entry points, wall ms inside them, ms of that blocked on the device."""
        o = np.zeros(3, np.float64)
        check(self._L.sb200_host_counters(self._h, ptr(o)))
        return {"calls": int(o[0]), "ms_total": float(o[1]), "ms_blocked": float(o[2])}

    def prefetch_inputs(self, boxes, features=None, has_feature=None, quality=None, custom_ids=None, own_area=None):
        """sb200_prefetch_inputs: start the H2D copy of a future request.  The arrays must be the very objects later
        passed to predict_batch (same memory) and C-contiguous with the right dtype (no conversion copies)."""
        for a, dt in ((boxes, np.float32), (features, np.float32), (has_feature, np.uint8), (quality, np.float32),
                      (custom_ids, np.int64), (own_area, np.float32)):
            if a is not None and not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags["C_CONTIGUOUS"]):
                raise ValueError("prefetch_inputs needs C-contiguous numpy arrays of the exact dtype")
        total = int(np.prod(boxes.shape)) // 6
        check(self._L.sb200_prefetch_inputs(self._h, total, ptr(boxes), ptr(features), ptr(has_feature), ptr(quality),
                                            ptr(custom_ids), ptr(own_area)))

    def predict_batch_device(self, scene_ids, det_offsets, d_boxes, d_features=0, d_has_feature=0, d_quality=0,
                             d_custom_ids=0, d_own_area=0, d_ids=0, d_epochs=0, d_lengths=0, d_voting_types=0,
                             d_predicted=0, d_observed=0):
        """Device-pointer call (sb200_predict_batch_device); d_* are raw device addresses (0 == NULL).  Stream-ordered:
        returns as soon as the frame is enqueued."""
        scene_ids = np.ascontiguousarray(scene_ids, dtype=np.uint64)
        det_offsets = np.ascontiguousarray(det_offsets, dtype=np.int32)
        vp = lambda a: C.c_void_p(a) if a else None  # noqa: E731
        po = PredictOut(vp(d_ids), vp(d_epochs), vp(d_lengths), vp(d_voting_types), vp(d_predicted), vp(d_observed))
        check(self._L.sb200_predict_batch_device(self._h, len(scene_ids), ptr(scene_ids), ptr(det_offsets), vp(d_boxes),
                                                 vp(d_features), vp(d_has_feature), vp(d_quality), vp(d_custom_ids),
                                                 vp(d_own_area), C.byref(po)))

    def skip_epochs(self, n, scene_id=0):
        check(self._L.sb200_skip_epochs(self._h, scene_id, n))

    def current_epoch(self, scene_id=0):
        return int(check(self._L.sb200_current_epoch(self._h, scene_id)))

    def active_tracks(self):
        return int(check(self._L.sb200_active_tracks(self._h)))

    def scene_track_counts(self, scene_ids):
        scene_ids = np.ascontiguousarray(scene_ids, dtype=np.uint64)
        out = np.zeros(len(scene_ids), np.int32)
        check(self._L.sb200_scene_track_counts(self._h, len(scene_ids), ptr(scene_ids), ptr(out)))
        return out

    def scene_live_counts(self, scene_ids):
        """(live tracks, feature blocks) per scene: what the device store holds / the visual cost kernel scans."""
        scene_ids = np.ascontiguousarray(scene_ids, dtype=np.uint64)
        live = np.zeros(len(scene_ids), np.int32)
        blocks = np.zeros(len(scene_ids), np.int32)
        check(self._L.sb200_scene_live_counts(self._h, len(scene_ids), ptr(scene_ids), ptr(live), ptr(blocks)))
        return live, blocks

    def set_auto_waste(self, periodicity):
        check(self._L.sb200_set_auto_waste(self._h, periodicity))

    def clear_wasted(self):
        check(self._L.sb200_clear_wasted(self._h))

    def wasted(self, cap=1 << 16):
        ids, sc = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
        ep, ln = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        pr, ob = np.zeros((cap, 6), np.float32), np.zeros((cap, 6), np.float32)
        n = check(self._L.sb200_wasted(self._h, cap, ptr(ids), ptr(sc), ptr(ep), ptr(ln), ptr(pr), ptr(ob)))
        return {"ids": ids[:n], "scene_ids": sc[:n], "epochs": ep[:n], "lengths": ln[:n], "predicted": pr[:n],
                "observed": ob[:n]}

    def wasted_history(self, cap=1 << 14, history_cap=None):
        """wasted() plus the box history of every wasted track (oldest first): predicted_history / observed_history are
        lists of [count][6] arrays."""
        H = int(history_cap if history_cap is not None else max(1, min(64, self.opts.history_length or 64)))
        ids, sc = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
        ep, ln = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        pr, ob = np.zeros((cap, 6), np.float32), np.zeros((cap, 6), np.float32)
        hp, ho = np.zeros((cap, H, 6), np.float32), np.zeros((cap, H, 6), np.float32)
        hc = np.zeros(cap, np.int32)
        n = check(self._L.sb200_wasted_history(self._h, cap, ptr(ids), ptr(sc), ptr(ep), ptr(ln), ptr(pr), ptr(ob), H,
                                               ptr(hp), ptr(ho), ptr(hc)))
        return {"ids": ids[:n], "scene_ids": sc[:n], "epochs": ep[:n], "lengths": ln[:n], "predicted": pr[:n],
                "observed": ob[:n], "predicted_history": [hp[i, : hc[i]].copy() for i in range(n)],
                "observed_history": [ho[i, : hc[i]].copy() for i in range(n)]}

    def idle_tracks(self, scene_id=0, cap=1 << 16):
        ids = np.zeros(cap, np.uint64)
        ep, ln = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        pr, ob = np.zeros((cap, 6), np.float32), np.zeros((cap, 6), np.float32)
        n = check(self._L.sb200_idle_tracks(self._h, scene_id, cap, ptr(ids), ptr(ep), ptr(ln), ptr(pr), ptr(ob)))
        return {"ids": ids[:n], "epochs": ep[:n], "lengths": ln[:n], "predicted": pr[:n], "observed": ob[:n]}

    def scene_tracks(self, scene_id=0, cap=1 << 14):
        ids = np.zeros(cap, np.uint64)
        bx, st = np.zeros((cap, 6), np.float32), np.zeros((cap, 30), np.float32)
        fc = np.zeros(cap, np.int32)
        n = check(self._L.sb200_scene_tracks(self._h, scene_id, cap, ptr(ids), ptr(bx), ptr(st), ptr(fc)))
        return {"ids": ids[:n], "boxes": bx[:n], "states": st[:n], "feat_counts": fc[:n]}

    def last_costs(self, scene_id=0, cap=1 << 22):
        out = np.zeros(cap, np.float32)
        m, n = C.c_int32(0), C.c_int32(0)
        cnt = check(self._L.sb200_last_costs(self._h, scene_id, cap, ptr(out), C.byref(m), C.byref(n)))
        return out[:cnt].reshape(m.value, n.value) if cnt else np.zeros((m.value, n.value), np.float32)

    def last_stage_ms(self):
        out = np.zeros(5, np.float32)
        check(self._L.sb200_last_stage_ms(self._h, ptr(out)))
        return dict(zip(("prep", "positional_cost", "visual_cost", "voting", "apply"), map(float, out)))


    def last_kernel_ms(self):
        out = np.zeros(2, np.float32)
        check(self._L.sb200_last_kernel_ms(self._h, ptr(out)))
        return {"vis_screen": float(out[0]), "vis_refine": float(out[1])}


class Comm:
    """NCCL communicator of the scene-sharded path (sb200_comm_*): scatter a request from an ingest rank to the ranks that
    own its scenes, gather the assigned track records back.  All pointers are raw device addresses."""

    def __init__(self, rank, world, unique_id: bytes, device):
        self._L = lib()
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        check(self._L.sb200_comm_create(rank, world, buf, device, C.byref(h)))
        self._h, self.rank, self.world = h, rank, world

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().sb200_comm_unique_id(buf))
        return buf.raw

    def close(self):
        if getattr(self, "_h", None):
            self._L.sb200_comm_destroy(self._h)
            self._h = None

    def scatter(self, root, det_range, feature_dim, all_boxes, all_features, my_boxes, my_features, stream):
        det_range = np.ascontiguousarray(det_range, dtype=np.int32)
        vp = lambda a: C.c_void_p(a) if a else None  # noqa: E731
        check(self._L.sb200_shard_scatter(self._h, root, ptr(det_range), feature_dim, vp(all_boxes), vp(all_features), None,
                                          None, None, vp(my_boxes), vp(my_features), None, None, None, C.c_void_p(stream)))

    def gather(self, root, det_range, mine: dict, all_: dict, stream):
        """mine / all_: {"ids": addr, "epochs": addr, "lengths": addr, "voting_types": addr} (device addresses)."""
        det_range = np.ascontiguousarray(det_range, dtype=np.int32)
        vp = lambda a: C.c_void_p(a) if a else None  # noqa: E731
        keys = ("ids", "epochs", "lengths", "voting_types", "predicted", "observed")
        pm = PredictOut(*[vp(mine.get(k, 0)) for k in keys])
        pa = PredictOut(*[vp((all_ or {}).get(k, 0)) for k in keys])
        check(self._L.sb200_shard_gather(self._h, root, ptr(det_range), C.byref(pm), C.byref(pa) if all_ else None,
                                         C.c_void_p(stream)))


def launch_count():
    """Kernels launched by libsimilari_b200.so since it was loaded."""
    return int(lib().sb200_launch_count())


# ------------------------------------------------------------------------------------------------ stateless operators
def sort_cost_matrix(positional_kind, cand_boxes, track_boxes, track_states30=None, iou_threshold=0.3,
                     min_confidence=0.05, pos_weight=1 / 20, vel_weight=1 / 160, device=0):
    cb, tb = _f32(cand_boxes).reshape(-1, 6), _f32(track_boxes).reshape(-1, 6)
    ts = _f32(track_states30).reshape(-1, 30) if track_states30 is not None else None
    out = np.empty((len(cb), len(tb)), np.float32)
    check(lib().sb200_sort_cost_matrix(positional_kind, iou_threshold, min_confidence, pos_weight, vel_weight, ptr(cb),
                                       len(cb), ptr(tb), ptr(ts), len(tb), ptr(out), device))
    return out


def visual_cost_matrix(visual_kind, threshold, cand_features, track_features, device=0):
    cf, tf = _f32(cand_features), _f32(track_features)
    out = np.empty((len(cf), len(tf)), np.float32)
    check(lib().sb200_visual_cost_matrix(visual_kind, threshold, ptr(cf), len(cf), ptr(tf), len(tf), cf.shape[1],
                                         ptr(out), device))
    return out


def sort_voting(threshold, cost_mn, device=0):
    c = _f32(cost_mn)
    w = np.full(c.shape[0], -1, np.int32)
    check(lib().sb200_sort_voting(threshold, ptr(c), c.shape[0], c.shape[1], ptr(w), device))
    return w


def visual_voting(positional_threshold, min_votes, pos_mn, vis_mnk, device=0):
    p, v = _f32(pos_mn), _f32(vis_mnk)
    m, n, k = v.shape
    w, vt = np.full(m, -1, np.int32), np.zeros(m, np.uint8)
    check(lib().sb200_visual_voting(positional_threshold, min_votes, ptr(p), ptr(v), m, n, k, ptr(w), ptr(vt), device))
    return w, vt


def kalman_initiate(boxes, pw=1 / 20, vw=1 / 160, device=0):
    b = _f32(boxes).reshape(-1, 6)
    out = np.empty((len(b), 30), np.float32)
    check(lib().sb200_kalman_initiate(pw, vw, ptr(b), len(b), ptr(out), device))
    return out


def kalman_predict(states30, pw=1 / 20, vw=1 / 160, device=0):
    s = _f32(states30).reshape(-1, 30)
    out = np.empty_like(s)
    check(lib().sb200_kalman_predict(pw, vw, ptr(s), len(s), ptr(out), device))
    return out


def kalman_update(states30, boxes, pw=1 / 20, vw=1 / 160, device=0):
    s, b = _f32(states30).reshape(-1, 30), _f32(boxes).reshape(-1, 6)
    out = np.empty_like(s)
    check(lib().sb200_kalman_update(pw, vw, ptr(s), ptr(b), len(s), ptr(out), device))
    return out


def own_area_shares(boxes, device=0):
    """exclusively_owned_areas_normalized_shares of ONE scene's boxes ([n][6]) on the GPU."""
    b = _f32(boxes).reshape(-1, 6)
    out = np.zeros(len(b), np.float32)
    check(lib().sb200_own_area_shares(ptr(b), len(b), ptr(out), device))
    return out


def nms_indices(boxes, scores, nms_threshold, score_threshold=None, device=0):
    b = _f32(boxes).reshape(-1, 6)
    s = _f32(scores) if scores is not None else None
    out = np.zeros(max(1, len(b)), np.int32)
    n = check(lib().sb200_nms(ptr(b), ptr(s), len(b), nms_threshold, 0.0 if score_threshold is None else score_threshold,
                              int(score_threshold is not None), ptr(out), device))
    return out[:n].copy()
