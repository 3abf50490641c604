"""Online sliding-window tracking with the scene kept on the device.

mot3d_b200.online.Tracker2D mirrors the reference's loop on Python objects (projects/PETS2009S2L1/
singleview_tracking.py:76-134, tracking_routines.py:16-52, scene.py:8-52): every batch it walks lists of Detection2D,
builds DetectionTracklet2D objects around the active trajectories and re-uploads all of it. DeviceTracker2D runs the
same loop on arrays: the active trajectories (frame, position, box per point, interpolated points included) stay in
device memory from one update() to the next, pass 1 leaves its tracks on the device, csrc/twopass.cu cuts them and
puts them next to the carried trajectories (k_split_tracks), csrc/tracklet.cu + the solver link them, and
k_online_concat writes the joined, hole-filled trajectories into the next state buffer. The host only decides the
bookkeeping the reference does in Scene.update / update_completed - which id a chain continues, which trajectories
are completed - from a few integers per tracklet (item, first frame, last frame) and the chains.

Scope: 2-D detections with boxes. Colour histograms are optional: with them, every point of the scene keeps its
histogram on the device (filled points have none) and the tracklet cost's appearance term compares the median
histograms of the windows (k_window_median_hist); without them a tracklet cost with use_color_histogram=True falls
back to the motion term alone, as the reference does when no tracklet has one.
valid_region (an object with is_inside(point), singleview_tracking.py:28-38) filters the detections of a batch on the
host before they are uploaded; the reference's second test, on the two ends of every new tracklet (:104-118), can
then never fail - every detection of a tracklet already passed the first - so nothing is skipped there."""
import ctypes

import numpy as np
import torch

from . import _lib
from .batch import DetectionBatch, _ptr, _stream
from .pipeline import TrackPipeline
from .spec import CostSpec, TrackletCostSpec
from .twopass import tracklet_pass, window_hist

__all__ = ["DeviceTracker2D"]


class DeviceTracker2D(object):
    """config: the `tracking` section Tracker2D takes (batch_size, detections{...}, tracklets{...});
    completed_after: Scene's. update(time_index, frame, pos, conf, bbox) takes the detections of one batch as arrays
    (frame int[n], pos float[n, 2], conf float[n], bbox float[n, 4], any order)."""

    def __init__(self, config, completed_after=30, device=None, split_length=10, th_length=2, valid_region=None):
        self.valid_region = valid_region
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        cfgd, cfgt = config["detections"], config["tracklets"]
        self.spec = CostSpec(kind="detections_2d", distance=cfgd["weight_distance"],
                             confidence=cfgd["weight_confidence"], weight_source_sink=cfgd["weight_source_sink"],
                             max_jump=cfgd["max_jump"])
        self.tspec = TrackletCostSpec(kind="tracklets_2d", distance=cfgt["weight_distance"],
                                      confidence=cfgt["weight_confidence"],
                                      weight_source_sink=cfgt["weight_source_sink"], max_jump=cfgt["max_jump"])
        self.window = int(cfgt["length_endpoints"]) if cfgt.get("length_endpoints") else 0
        self.split_length, self.th_length = int(split_length), int(th_length)
        self.completed_after = completed_after
        self.id_last = 0
        self.n_tracklets_seen = 0   # Scene.id_last_tracklet
        self.ids = []               # active ids, ascending = the order of Scene.active
        self.ptr = np.zeros(1, dtype=np.int64)   # host copy of the point offsets of the active trajectories
        self.completed = {}         # id -> (frame[int32], pos[npts, 2], bbox[npts, 4]) host arrays
        dev = self.device
        self._cap = 0
        self._state = None          # dict(ptr=int32 device [A+1], frame, pos, bbox) of the active trajectories
        self._spare = None
        self._pipe = None
        self._hist_shape = None     # (C, B) once a batch came with histograms
        self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        self._zero = torch.zeros(2, dtype=torch.int32, device=dev)

    def _pipeline(self, n):
        """The preallocated, synchronisation-free detection pass (pipeline.TrackPipeline), grown when a batch is larger."""
        if self._pipe is None or self._pipe.n_max < n:
            n_max = max(1024, 2 * n)
            self._pipe = TrackPipeline(n_max, 1, 64 * n_max, device=self.device, dim=2)
        return self._pipe

    # ---- state buffers --------------------------------------------------------------------------------------------
    def _buffers(self, points):
        """A state buffer (not the live one) with room for `points` points."""
        dev = self.device
        b = self._spare
        if b is None or b["cap"] < points or (self._hist_shape is not None and b["hist"] is None):
            cap = max(1024, int(points * 1.5))
            b = dict(cap=cap, frame=torch.empty(cap, dtype=torch.int32, device=dev),
                     pos=torch.empty(2 * cap, dtype=torch.float64, device=dev),
                     bbox=torch.empty(4 * cap, dtype=torch.float64, device=dev), ptr=None, hist=None, has=None)
            if self._hist_shape is not None:
                b["hist"] = torch.empty((cap,) + self._hist_shape, dtype=torch.float64, device=dev)
                b["has"] = torch.empty(cap, dtype=torch.uint8, device=dev)
        self._spare = None
        return b

    # ---- one batch ------------------------------------------------------------------------------------------------
    def update(self, time_index, frame, pos, conf, bbox, hist=None):
        """hist: None or float[n, C, B], the colour histograms of the detections (mot3d_b200.color_histograms writes
        that layout); every batch of a run must come with or without them."""
        L = _lib.lib()
        dev = self.device
        st = _stream(dev)
        i32 = dict(dtype=torch.int32, device=dev)
        f64 = dict(dtype=torch.float64, device=dev)
        if self.valid_region is not None and len(frame):
            inside = np.fromiter((bool(self.valid_region.is_inside(p)) for p in np.asarray(pos)), dtype=bool,
                                 count=len(frame))
            frame, pos, conf, bbox = (np.asarray(a)[inside] for a in (frame, pos, conf, bbox))
            hist = None if hist is None else np.asarray(hist)[inside]
        if hist is not None:
            shape = tuple(np.shape(hist)[1:])
            if len(shape) != 2 or (self._hist_shape not in (None, shape)) or (self._hist_shape is None and self.id_last):
                raise ValueError("histograms must be [n, C, B] with the same C, B in every batch of a run")
            self._hist_shape = shape
        elif self._hist_shape is not None and len(frame):
            raise ValueError("earlier batches came with histograms, this one does not")
        n = int(len(frame))
        A = len(self.ids)
        P = int(self.ptr[-1])
        # (1) pass 1: the tracks of this batch (compute_trajectories, tracking_routines.py:16-34)
        if n > 0:
            hb, _ = DetectionBatch.from_arrays(np.array([0, n]), frame, pos, conf, bbox=bbox, hist=hist)
            db = hb.to(dev)
            pipe = self._pipeline(n)
            pipe.run(db, self.spec)   # no host synchronisation; its flag word is read with the tracklet count below
            t_count, t_ptr, t_det = pipe.track_count, pipe.track_ptr, pipe.track_det
            b_frame, b_pos, b_bbox, b_hist = db.frame, db.pos, db.bbox, db.hist
        else:
            t_count, t_ptr, t_det = self._zero[:1], self._zero, self._zero
            b_frame = b_pos = b_bbox = b_hist = None
        if A == 0 and n == 0:
            return 0
        # (2) the tracklet set: carried trajectories + pieces of the new tracks (singleview_tracking.py:94-121)
        m_max = A + n
        stride = P + n + 1
        n_trk = torch.zeros(1, **i32)
        item, trk_first, trk_len, key = (torch.empty(m_max + 1, **i32) for _ in range(4))
        span = torch.empty(2 * m_max + 2, **i32)
        head_ptr, tail_ptr = torch.empty(m_max + 2, **i32), torch.empty(m_max + 2, **i32)
        head_idx, tail_idx = torch.empty(stride, **i32), torch.empty(stride, **i32)
        head_pos, tail_pos = torch.empty(2 * stride, **f64), torch.empty(2 * stride, **f64)
        wsb = L.mot3d_tracklets_from_tracks_workspace_bytes(n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        s = self._state
        _lib.check(L.mot3d_online_tracklets(
            n, 2, _ptr(b_frame), _ptr(b_pos), _ptr(t_count), _ptr(t_ptr), _ptr(t_det), self.split_length,
            self.th_length, self.window, A, _ptr(s["ptr"]) if A else None, _ptr(s["frame"]) if A else None,
            _ptr(s["pos"]) if A else None, s["cap"] if A else 0, stride, _ptr(n_trk), _ptr(item), _ptr(trk_first),
            _ptr(trk_len), _ptr(key), _ptr(head_ptr), _ptr(tail_ptr), _ptr(head_idx), _ptr(tail_idx), _ptr(head_pos),
            _ptr(tail_pos), _ptr(span), _ptr(ws), wsb, st))
        if n > 0:
            M, flag = torch.cat([n_trk, pipe.overflow]).tolist()
            if flag:
                if flag & (_lib.ERR_ARC_CAPACITY | _lib.ERR_WINDOW):
                    # the arc buffer was a guess: size it from what the build counted and run the batch again
                    self._pipe = TrackPipeline(pipe.n_max, 1, 2 * pipe.n_edges() + 1024, device=dev, dim=2)
                    return self.update(time_index, frame, pos, conf, bbox, hist)
                self._pipe = None   # (its flag word is sticky: the next batch gets fresh buffers)
                pipe.check_overflow()   # raises: frames out of the key range
        else:
            M = int(n_trk.item())
        self.n_tracklets_seen += M - A
        if M == 0:
            return 0
        # (3) pass 2 (concatenate_tracklets, tracking_routines.py:36-52)
        wh = None
        if self._hist_shape is not None and self.tspec.use_hist:
            wh = window_hist(dev, M, self._hist_shape, self.window, item, trk_first, trk_len, t_det, b_frame, b_hist, A,
                             s["frame"] if A else None, s["hist"] if A else None, s["has"] if A else None)
        c_count, c_ptr, c_trk, _total, _E = tracklet_pass(dev, 2, M, n_trk, stride, head_ptr, tail_ptr, head_idx,
                                                          tail_idx, head_pos, tail_pos, key, self.tspec, hist=wh)
        meta = torch.cat([c_count, c_ptr[:M + 1], c_trk[:M], item[:M], span[:2 * M]]).cpu().numpy()
        K = int(meta[0])
        c_ptr_h = meta[1:M + 2]
        c_trk_h = meta[M + 2:2 * M + 2]
        item_h = meta[2 * M + 2:3 * M + 2]
        span_h = meta[3 * M + 2:5 * M + 2].reshape(M, 2)
        # (4) Scene.update for every chain, in the solver's order (singleview_tracking.py:123-131, scene.py:24-31):
        #     a chain that starts with a carried trajectory replaces it, any other chain is a new trajectory; carried
        #     trajectories the solver left alone, or used in the middle of a chain, stay as they are
        rank_of = np.empty(M, dtype=np.int64)
        rank_of[item_h] = np.arange(M)
        chains = {self.ids[c]: [int(rank_of[c])] for c in range(A)}
        for k in range(K):
            ranks = c_trk_h[c_ptr_h[k]:c_ptr_h[k + 1]].tolist()
            first = int(item_h[ranks[0]])
            if first < A:
                chains[self.ids[first]] = ranks
            else:
                chains[self.id_last] = ranks
                self.id_last += 1
        # (5) Scene.update_completed (scene.py:33-38); layout of the next state: active ids first, then the completed
        last_of = {i: int(span_h[r[-1], 1]) for i, r in chains.items()}
        keep = [i for i in chains if not (last_of[i] + self.completed_after < time_index)]
        done = [i for i in chains if last_of[i] + self.completed_after < time_index]
        order = keep + done
        npts = [last_of[i] - int(span_h[chains[i][0], 0]) + 1 for i in order]
        out_ptr = np.concatenate([[0], np.cumsum(npts)]).astype(np.int32)
        ch_len = [len(chains[i]) for i in order]
        ch_ptr = np.concatenate([[0], np.cumsum(ch_len)]).astype(np.int32)
        ch = np.fromiter((r for i in order for r in chains[i]), dtype=np.int32, count=int(ch_ptr[-1]))
        n_out = len(order)
        up = torch.from_numpy(np.concatenate([out_ptr, ch_ptr, ch])).to(dev)
        d_out_ptr, d_ch_ptr, d_ch = up[:n_out + 1], up[n_out + 1:2 * n_out + 2], up[2 * n_out + 2:]
        total = int(out_ptr[-1])
        nb = self._buffers(total)
        _lib.check(L.mot3d_online_concat(
            n_out, _ptr(d_ch_ptr), _ptr(d_ch), _ptr(d_out_ptr), _ptr(item), _ptr(trk_first), _ptr(trk_len), A, n, 2,
            _ptr(t_det), _ptr(b_frame), _ptr(b_pos), _ptr(b_bbox), _ptr(s["frame"]) if A else None,
            _ptr(s["pos"]) if A else None, _ptr(s["bbox"]) if A else None, s["cap"] if A else 0, _ptr(nb["frame"]),
            _ptr(nb["pos"]), _ptr(nb["bbox"]), nb["cap"],
            0 if self._hist_shape is None else self._hist_shape[0] * self._hist_shape[1], _ptr(b_hist),
            _ptr(s["hist"]) if A else None, _ptr(s["has"]) if A else None, _ptr(nb["hist"]), _ptr(nb["has"]),
            _ptr(self._err), st))
        nb["ptr"] = d_out_ptr
        # completed trajectories leave the device
        if done:
            lo, hi = int(out_ptr[len(keep)]), total
            fr = nb["frame"][lo:hi].cpu().numpy()
            ps = torch.stack([nb["pos"][c * nb["cap"] + lo:c * nb["cap"] + hi] for c in range(2)], 1).cpu().numpy()
            bx = torch.stack([nb["bbox"][c * nb["cap"] + lo:c * nb["cap"] + hi] for c in range(4)], 1).cpu().numpy()
            for k, i in enumerate(done):
                a, b = int(out_ptr[len(keep) + k]) - lo, int(out_ptr[len(keep) + k + 1]) - lo
                self.completed[i] = (fr[a:b], ps[a:b], bx[a:b])
        if int(self._err.item()):
            raise _lib.Mot3dError(_lib.E_INVALID, "online write-back: a chain of tracklets does not cover its frame range")
        self._spare, self._state = self._state, nb
        self.ids = keep
        self.ptr = out_ptr[:len(keep) + 1].astype(np.int64)
        return 0

    # ---- reading the scene ----------------------------------------------------------------------------------------
    def active(self):
        """id -> (frame[npts], pos[npts, 2], bbox[npts, 4]) of the active trajectories (device -> host copy)."""
        if not self.ids:
            return {}
        s, P = self._state, int(self.ptr[-1])
        fr = s["frame"][:P].cpu().numpy()
        ps = torch.stack([s["pos"][c * s["cap"]:c * s["cap"] + P] for c in range(2)], 1).cpu().numpy()
        bx = torch.stack([s["bbox"][c * s["cap"]:c * s["cap"] + P] for c in range(4)], 1).cpu().numpy()
        return {i: (fr[self.ptr[k]:self.ptr[k + 1]], ps[self.ptr[k]:self.ptr[k + 1]], bx[self.ptr[k]:self.ptr[k + 1]])
                for k, i in enumerate(self.ids)}
