"""Tracklet linking (the reference's second pass): build_graph on DetectionTracklet2D/3D lists.

Packing (host) -> CUDA arc kernels (csrc/tracklet.cu: per-end line fits, all-pairs costs, warp-ballot compaction in
the reference's arc order) -> the same solver / track extraction as for detections. Tracklets are not frame sorted
and overlap in time (reference mot3d/mot3d.py:134-135), so the solver's layer order comes from the arcs
(mot3d_b200/dimacs.py); these graphs are small (6-310 nodes in the reference's configs)."""
import ctypes

import numpy as np
import torch

from . import _lib
from .batch import _ptr, _stream, solve_graphs, extract_tracks, tracks_to_lists
from .types import DetectionTracklet3D


def pack_tracklets(tracklets, spec):
    """DetectionTracklet objects -> dict of host arrays (layout of include/mot3d_b200.h: mot3d_tracklets)."""
    dim = spec.dim
    n = len(tracklets)
    hp, tp = [0], [0]
    hidx, tidx, hpos, tpos = [], [], [], []
    for t in tracklets:
        hidx.extend(int(i) for i in t.head.indexes)
        hpos.extend(np.asarray(q, dtype=np.float64).reshape(-1)[:dim] for q in t.head.positions)
        hp.append(len(hidx))
        tidx.extend(int(i) for i in t.tail.indexes)
        tpos.extend(np.asarray(q, dtype=np.float64).reshape(-1)[:dim] for q in t.tail.positions)
        tp.append(len(tidx))
    out = dict(n=n, dim=dim, head_ptr=np.array(hp, dtype=np.int32), tail_ptr=np.array(tp, dtype=np.int32),
               head_idx=np.array(hidx, dtype=np.int32), tail_idx=np.array(tidx, dtype=np.int32),
               head_pos=np.ascontiguousarray(np.array(hpos, dtype=np.float64).reshape(-1, dim).T),
               tail_pos=np.ascontiguousarray(np.array(tpos, dtype=np.float64).reshape(-1, dim).T),
               conf=np.array([t.confidence for t in tracklets], dtype=np.float64),
               tail_access=np.array([bool(getattr(t.tail, "access_point", False)) for t in tracklets], dtype=np.uint8))
    flags = None
    if any((not t.entry_points) or (not t.exit_point) for t in tracklets):
        flags = np.array([(_lib.FLAG_ENTRY if t.entry_points else 0) | (_lib.FLAG_EXIT if t.exit_point else 0)
                          for t in tracklets], dtype=np.uint8)
    out["flags"] = flags
    if spec.use_hist:
        if isinstance(tracklets[0], DetectionTracklet3D):
            # per-view 2-D tracklets, in each tracklet's dict order (reference mot3d/weight_functions.py:146-158)
            labels = {}
            ptr, ids, vh, vt = [0], [], [], []
            for t in tracklets:
                for view, t2 in t.tracklets_2d.items():
                    h1, h2 = t2.head.color_histogram, t2.tail.color_histogram
                    if h1 is None or h2 is None:
                        raise ValueError("use_color_histogram=True but the 2D tracklet of view %r has no "
                                         "color_histogram" % (view,))
                    ids.append(labels.setdefault(view, len(labels)))
                    vh.append(np.asarray(h1, dtype=np.float64))
                    vt.append(np.asarray(h2, dtype=np.float64))
                ptr.append(len(ids))
            out["head_hist"] = out["tail_hist"] = None
            out["view_ptr"] = np.array(ptr, dtype=np.int32)
            out["view_id"] = np.array(ids, dtype=np.int32)
            out["view_head_hist"] = np.array(vh, dtype=np.float64)
            out["view_tail_hist"] = np.array(vt, dtype=np.float64)
            return out
        hh = [getattr(t.head, "color_histogram", None) for t in tracklets]
        th = [getattr(t.tail, "color_histogram", None) for t in tracklets]
        shape = next((np.asarray(h).shape for h in hh + th if h is not None), None)
        if shape is None:
            out["head_hist"] = out["tail_hist"] = None  # nobody has a histogram: the motion term alone decides
        else:
            z = np.zeros(shape)
            out["head_hist"] = np.array([z if h is None else np.asarray(h, dtype=np.float64) for h in hh])
            out["tail_hist"] = np.array([z if h is None else np.asarray(h, dtype=np.float64) for h in th])
            out["head_has_hist"] = np.array([h is not None for h in hh], dtype=np.uint8)
            out["tail_has_hist"] = np.array([h is not None for h in th], dtype=np.uint8)
    return out


def build_tracklet_arcs(packed, spec, device=None):
    """Arc build on the GPU. Returns host arrays (row_ptr, col, weight f64, cost i64), rows = tail tracklets, columns
    ascending = the reference's arc order."""
    L = _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n, dim = packed["n"], packed["dim"]

    def up(a):
        return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    t_ = {k: up(packed.get(k)) for k in ("head_ptr", "tail_ptr", "head_idx", "tail_idx", "head_pos", "tail_pos",
                                          "tail_access", "head_hist", "tail_hist", "head_has_hist", "tail_has_hist",
                                          "view_ptr", "view_id", "view_head_hist", "view_tail_hist")}
    cs = spec.to_c()
    view_mode = spec.use_hist and t_["view_ptr"] is not None and int(t_["view_id"].shape[0]) > 0
    use_hist = spec.use_hist and (t_["head_hist"] is not None or view_mode)
    cs.use_hist = int(use_hist)
    fit_h = torch.empty(max(n, 1) * dim * 2, dtype=torch.float64, device=dev)
    fit_t = torch.empty(max(n, 1) * dim * 2, dtype=torch.float64, device=dev)
    trk = _lib.Tracklets()
    trk.n, trk.dim = n, dim
    trk.n_head_points, trk.n_tail_points = int(packed["head_ptr"][-1]), int(packed["tail_ptr"][-1])
    for name in ("head_ptr", "tail_ptr", "head_idx", "tail_idx", "head_pos", "tail_pos"):
        setattr(trk, name, t_[name].data_ptr())
    trk.tail_access_point = t_["tail_access"].data_ptr()
    if view_mode:
        trk.hist_channels, trk.hist_bins = int(t_["view_head_hist"].shape[1]), int(t_["view_head_hist"].shape[2])
        trk.view_ptr, trk.view_id = t_["view_ptr"].data_ptr(), t_["view_id"].data_ptr()
        trk.view_head_hist, trk.view_tail_hist = t_["view_head_hist"].data_ptr(), t_["view_tail_hist"].data_ptr()
    elif use_hist:
        trk.hist_channels, trk.hist_bins = int(t_["head_hist"].shape[1]), int(t_["head_hist"].shape[2])
        trk.head_hist, trk.tail_hist = t_["head_hist"].data_ptr(), t_["tail_hist"].data_ptr()
        trk.head_has_hist, trk.tail_has_hist = t_["head_has_hist"].data_ptr(), t_["tail_has_hist"].data_ptr()
    trk.head_fit_ws, trk.tail_fit_ws = fit_h.data_ptr(), fit_t.data_ptr()
    st = _stream(dev)
    row_count = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    _lib.check(L.mot3d_tracklet_edge_count(ctypes.byref(trk), ctypes.byref(cs), _ptr(row_count), st))
    row_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    wsb = L.mot3d_scan_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    _lib.check(L.mot3d_exclusive_scan_i32(_ptr(row_count), n, _ptr(row_ptr), _ptr(ws), wsb, st))
    e = int(row_ptr[n].item())
    col = torch.empty(max(e, 1), dtype=torch.int32, device=dev)
    cost = torch.empty(max(e, 1), dtype=torch.int64, device=dev)
    weight = torch.empty(max(e, 1), dtype=torch.float64, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(L.mot3d_tracklet_edge_fill(ctypes.byref(trk), ctypes.byref(cs), _ptr(row_ptr), e, _ptr(col), _ptr(cost),
                                          _ptr(weight), _ptr(overflow), st))
    flag = int(overflow.item())
    if flag & 2:
        raise ValueError("two linked 3-D tracklets share no 2-D view: weight_distance_tracklets_3d is nan for them "
                         "(np.mean of an empty list, reference mot3d/weight_functions.py:146-158)")
    return (row_ptr.cpu().numpy().astype(np.int64), col[:e].cpu().numpy().astype(np.int64), weight[:e].cpu().numpy(),
            cost[:e].cpu().numpy())


def solve_tracklet_graph(n, row_ptr, col, cost, entry_cost, node_cost, exit_cost):
    """Arcs (out-CSR over tracklets in input order) + per-node costs -> (tracks as lists of tracklet indices in the
    reference's order, total cost, K). The solve and the track extraction run on the GPU."""
    from . import dimacs
    m = np.arange(n)
    deg = np.diff(row_ptr)
    has_en, has_ex = entry_cost < _lib.INF_COST, exit_cost < _lib.INF_COST
    tail = np.concatenate([np.ones(int(has_en.sum()), dtype=np.int64), 2 + 2 * m, (3 + 2 * m)[has_ex],
                           np.repeat(3 + 2 * m, deg)])
    head = np.concatenate([(2 + 2 * m)[has_en], 3 + 2 * m, np.full(int(has_ex.sum()), 2 * n + 2, dtype=np.int64),
                           2 + 2 * col])
    c = np.concatenate([entry_cost[has_en], node_cost, exit_cost[has_ex], cost])
    gb, perm = dimacs.graph_from_arcs(2 * n + 2, tail, head, c)
    sol = solve_graphs(gb)
    extract_tracks(gb, sol)
    tracks = tracks_to_lists(np.array([0, n]), sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(),
                             sol.track_det.cpu().numpy())[0]
    tracks = [[int(perm[k]) for k in t] for t in tracks]
    tracks.sort(key=lambda t: t[0])  # the reference enumerates S's successors in node order (mot3d/mot3d.py:175)
    return tracks, int(sol.total_cost[0].item()), int(sol.n_paths[0].item()), sol
