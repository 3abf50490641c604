"""Detection pass -> tracklet pass without leaving the device.

The reference runs its second pass on Python objects: every trajectory of pass 1 is cut by split_trajectory_modulo,
short pieces are dropped, DetectionTracklet2D/3D objects are built around the pieces, build_graph / solve_graph run on
those, and concat_tracklets joins the linked pieces (examples/soldiers_3d/example.py:66-104,
mot3d/utils/trajectory.py:11-12,216-248, mot3d/types.py:94-184). link_tracks does the same on the arrays pass 1 left
on the GPU: csrc/twopass.cu cuts and windows the tracks, csrc/tracklet.cu builds the arcs, the solver and the track
extraction are the ones of pass 1, and only the final trajectories (indices of pass-1 detections) come back. Between
the passes the host reads two integers (the number of tracklets and the number of arcs, to size the launches).

Scope: one graph per call. 2-D tracklet costs with the appearance term use the median histograms of the windows
(k_window_median_hist) when the batch carries histograms; 3-D tracklets with use_color_histogram=True (per-view 2-D
tracklets) take the object path."""
import ctypes

import numpy as np
import torch

from . import _lib
from .batch import _ptr, _stream
from .spec import quantise_scalar


def tracklet_pass(dev, dim, M, n_trk, stride, head_ptr, tail_ptr, head_idx, tail_idx, head_pos, tail_pos, key, tspec,
                  hist=None):
    """Arcs between M tracklets (windows in the mot3d_tracklets arrays, planar stride `stride`), solve, chains.
    hist: None, or (head_hist [M, C, B], tail_hist, head_has uint8[M], tail_has) from window_hist() - the appearance
    term of the tracklet cost then counts wherever both ends have a histogram.
    Returns (t_count, t_ptr, t_trk, total_cost, n_arcs): device tensors except n_arcs."""
    L = _lib.lib()
    st = _stream(dev)
    i32 = dict(dtype=torch.int32, device=dev)
    i64 = dict(dtype=torch.int64, device=dev)
    f64 = dict(dtype=torch.float64, device=dev)
    n = stride
    # ---- arcs between tracklets (csrc/tracklet.cu), rows = arc tails
    trk = _lib.Tracklets()
    trk.n, trk.dim = M, dim
    trk.n_head_points = trk.n_tail_points = n  # planar stride of head_pos / tail_pos
    trk.head_ptr, trk.tail_ptr = head_ptr.data_ptr(), tail_ptr.data_ptr()
    trk.head_idx, trk.tail_idx = head_idx.data_ptr(), tail_idx.data_ptr()
    trk.head_pos, trk.tail_pos = head_pos.data_ptr(), tail_pos.data_ptr()
    access = torch.zeros(M, dtype=torch.uint8, device=dev)  # tracklets built from solved tracks: no access points
    trk.tail_access_point = access.data_ptr()
    fit_h, fit_t = torch.empty(M * dim * 2, **f64), torch.empty(M * dim * 2, **f64)
    trk.head_fit_ws, trk.tail_fit_ws = fit_h.data_ptr(), fit_t.data_ptr()
    cs = tspec.to_c()
    cs.use_hist = 0
    if hist is not None and tspec.use_hist:
        cs.use_hist = 1
        trk.hist_channels, trk.hist_bins = int(hist[0].shape[1]), int(hist[0].shape[2])
        trk.head_hist, trk.tail_hist = hist[0].data_ptr(), hist[1].data_ptr()
        trk.head_has_hist, trk.tail_has_hist = hist[2].data_ptr(), hist[3].data_ptr()
    row_count = torch.zeros(M, **i32)
    _lib.check(L.mot3d_tracklet_edge_count(ctypes.byref(trk), ctypes.byref(cs), _ptr(row_count), st))
    row_ptr = torch.empty(M + 1, **i32)
    sb = L.mot3d_scan_workspace_bytes(M)
    sws = torch.empty(sb, dtype=torch.uint8, device=dev)
    _lib.check(L.mot3d_exclusive_scan_i32(_ptr(row_count), M, _ptr(row_ptr), _ptr(sws), sb, st))
    E = int(row_ptr[M].item())
    col, cost = torch.empty(max(E, 1), **i32), torch.empty(max(E, 1), **i64)
    weight = torch.empty(max(E, 1), **f64)
    overflow = torch.zeros(1, **i32)
    _lib.check(L.mot3d_tracklet_edge_fill(ctypes.byref(trk), ctypes.byref(cs), _ptr(row_ptr), E, _ptr(col), _ptr(cost),
                                          _ptr(weight), _ptr(overflow), st))
    # ---- the solver's inputs: in-CSR, constant node costs (DetectionTracklet's default confidence 0.5,
    #      mot3d/types.py:49), time key = first frame of every tracklet
    in_ptr, in_src, in_cost = torch.empty(M + 1, **i32), torch.empty(max(E, 1), **i32), torch.empty(max(E, 1), **i64)
    en, cn, ex = torch.empty(M, **i64), torch.empty(M, **i64), torch.empty(M, **i64)
    gptr = torch.empty(2, **i32)
    cur = torch.empty(M + 1, **i32)
    conf = 0.5
    _lib.check(L.mot3d_tracklet_graph_in_csr(
        _ptr(n_trk), M, _ptr(row_ptr), _ptr(col), _ptr(cost), quantise_scalar(tspec.weight_source_sink),
        quantise_scalar(-(conf * tspec.confidence["mul"] + tspec.confidence["bias"])), 0, _ptr(in_ptr), _ptr(in_src),
        _ptr(in_cost), _ptr(en), _ptr(cn), _ptr(ex), _ptr(gptr), _ptr(cur), st))
    nxt, prv = torch.empty(M, **i32), torch.empty(M, **i32)
    n_paths, total = torch.zeros(1, **i32), torch.zeros(1, **i64)
    swb = L.mot3d_solve_workspace_bytes(M, 1, E)
    sw = torch.empty(swb, dtype=torch.uint8, device=dev)
    _lib.check(L.mot3d_solve_batch(1, M, _ptr(gptr), M, _ptr(key), E, _ptr(in_ptr), _ptr(in_src), _ptr(in_cost), _ptr(en),
                                   _ptr(cn), _ptr(ex), _ptr(nxt), _ptr(prv), _ptr(n_paths), _ptr(total), None, None,
                                   _ptr(sw), swb, st))
    t_count, t_ptr, t_trk = torch.zeros(1, **i32), torch.zeros(M + 2, **i32), torch.zeros(M, **i32)
    _lib.check(L.mot3d_extract_tracks(1, _ptr(gptr), _ptr(nxt), _ptr(prv), _ptr(t_count), _ptr(t_ptr), _ptr(t_trk), st))
    return t_count, t_ptr, t_trk, total, E


def window_hist(dev, M, shape, window, item, trk_first, trk_len, track_det, frame, hist, n_carry=0, carry_frame=None,
                carry_hist=None, carry_has=None):
    """Median histograms of the head / tail windows of M tracklets (csrc/twopass.cu k_window_median_hist).
    shape = (C, B). Returns the tuple tracklet_pass(hist=...) takes."""
    L = _lib.lib()
    C, B = shape
    hh, th = (torch.empty((M, C, B), dtype=torch.float64, device=dev) for _ in range(2))
    hs, ts = (torch.empty(M, dtype=torch.uint8, device=dev) for _ in range(2))
    _lib.check(L.mot3d_tracklet_window_hist(M, _ptr(item), _ptr(trk_first), _ptr(trk_len), n_carry,
                                            0 if window is None else int(window), C * B, _ptr(track_det), _ptr(frame),
                                            _ptr(hist), _ptr(carry_frame), _ptr(carry_hist), _ptr(carry_has), _ptr(hh),
                                            _ptr(th), _ptr(hs), _ptr(ts), _stream(dev)))
    return hh, th, hs, ts


def link_tracks(batch, sol, tspec, split_length=10, th_length=10, window=None):
    """batch: DetectionBatch on a CUDA device holding ONE graph; sol: its Solution after extract_tracks;
    tspec: TrackletCostSpec (kind, distance kwargs, confidence kwargs, weight_source_sink, max_jump) of pass 2.
    Returns (trajectories, info): trajectories = list of lists of detection indices (positions in `batch`), ordered by
    their first detection; info = dict(n_tracklets, n_arcs, total_cost (1e-7 units), tracklets = (first, len, key)
    device tensors)."""
    L = _lib.lib()
    if batch.n_graphs != 1:
        raise ValueError("link_tracks takes one graph per call")
    if tspec.use_hist and tspec.kind != "tracklets_2d":
        raise NotImplementedError("link_tracks: 3-D tracklet costs with use_color_histogram=True average over the "
                                  "per-view 2-D tracklets (host path: build_graph on DetectionTracklet3D objects)")
    dev = batch.frame.device
    n, dim = batch.n, batch.dim
    st = _stream(dev)
    i32 = dict(dtype=torch.int32, device=dev)
    i64 = dict(dtype=torch.int64, device=dev)
    f64 = dict(dtype=torch.float64, device=dev)
    n_trk = torch.zeros(1, **i32)
    trk_first, trk_len, key = torch.empty(n + 1, **i32), torch.empty(n + 1, **i32), torch.empty(n + 1, **i32)
    head_ptr, tail_ptr = torch.empty(n + 2, **i32), torch.empty(n + 2, **i32)
    head_idx, tail_idx = torch.empty(n + 1, **i32), torch.empty(n + 1, **i32)
    head_pos, tail_pos = torch.empty(dim * n + 1, **f64), torch.empty(dim * n + 1, **f64)
    wsb = L.mot3d_tracklets_from_tracks_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    _lib.check(L.mot3d_tracklets_from_tracks(
        n, dim, _ptr(batch.frame), _ptr(batch.pos), _ptr(sol.track_count), _ptr(sol.track_ptr), _ptr(sol.track_det),
        int(split_length), int(th_length), 0 if window is None else int(window), _ptr(n_trk), _ptr(trk_first),
        _ptr(trk_len), _ptr(key), _ptr(head_ptr), _ptr(tail_ptr), _ptr(head_idx), _ptr(tail_idx), _ptr(head_pos),
        _ptr(tail_pos), _ptr(ws), wsb, st))
    M = int(n_trk.item())
    info = dict(n_tracklets=M, n_arcs=0, total_cost=0, tracklets=(trk_first[:M], trk_len[:M], key[:M]))
    if M == 0:
        return [], info
    hist = None
    if tspec.use_hist and batch.hist is not None:
        hist = window_hist(dev, M, tuple(batch.hist.shape[1:]), window, None, trk_first, trk_len, sol.track_det,
                           batch.frame, batch.hist)
    t_count, t_ptr, t_trk, total, E = tracklet_pass(dev, dim, M, n_trk, n, head_ptr, tail_ptr, head_idx, tail_idx,
                                                    head_pos, tail_pos, key, tspec, hist=hist)
    info["n_arcs"] = E
    # ---- chains of tracklets -> chains of detections
    o_count, o_ptr, o_det = torch.zeros(1, **i32), torch.zeros(M + 2, **i32), torch.zeros(n + 1, **i32)
    _lib.check(L.mot3d_concat_tracklets(_ptr(t_count), _ptr(t_ptr), _ptr(t_trk), _ptr(trk_first), _ptr(trk_len),
                                        _ptr(sol.track_det), _ptr(o_count), _ptr(o_ptr), _ptr(o_det), st))
    K = int(o_count.item())
    ptr = o_ptr[:K + 1].cpu().numpy()
    det = o_det[:int(ptr[K]) if K else 0].cpu().numpy()
    info["total_cost"] = int(total.item())
    info["chains"] = (t_count, t_ptr, t_trk)
    trajectories = [det[int(ptr[k]):int(ptr[k + 1])].tolist() for k in range(K)]
    return trajectories, info
