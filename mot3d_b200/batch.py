"""Batched tracking on the GPU: many independent graphs (windows / camera views / sequences) per call.

DetectionBatch  packed detections (structure of arrays, torch tensors) for G graphs
build_graphs    frame sort + single-pass x-window arc build (or count -> scan -> fill)  (reference mot3d/mot3d.py:75-141)
solve_graphs    components -> first tree -> single-track pass -> SSP per component    (reference wrapper.cpp:171-317)
extract_tracks  flow-following kernel                                        (reference mot3d/mot3d.py:165-184)
track_batch     the three in a row; track_batch_host: the same from pinned host buffers through one C-ABI call

Everything here is plumbing around include/mot3d_b200.h: torch allocates, ctypes passes pointers.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .spec import CostSpec


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class DetectionBatch(object):
    """G graphs, detections of graph g in [graph_ptr[g], graph_ptr[g+1]), each graph stably sorted by frame
    (the reference's node order, mot3d/mot3d.py:61-64). Arrays may be numpy (host) or torch tensors.

    frame int32[N]; pos float64[dim, N] (planar); conf float64[N]; flags uint8[N] or None;
    bbox float64[4, N] or None; hist float64[N, C, B] or None."""

    def __init__(self, graph_ptr, frame, pos, conf, flags=None, bbox=None, hist=None, views=None):
        self.views = views  # 3-D detections: dict(view_ptr int32[N+1], view_id int32[E], bbox float64[4, E] or None,
        #                     hist float64[E, C, B] or None) - Detection3D.detections_2d, entries in dict order
        self.graph_ptr = graph_ptr
        self.frame = frame
        self.pos = pos
        self.conf = conf
        self.flags = flags
        self.bbox = bbox
        self.hist = hist

    @property
    def n(self):
        return int(self.frame.shape[0])

    @property
    def n_graphs(self):
        return int(self.graph_ptr.shape[0]) - 1

    @property
    def dim(self):
        return int(self.pos.shape[0])

    @staticmethod
    def from_arrays(graph_ptr, frame, pos, conf=None, flags=None, bbox=None, hist=None, sort=True):
        """pos: [N, dim] (as users hold it). Sorts each graph stably by frame unless sort=False.
        Returns (batch, order) where order[k] = index in the input of the k-th packed detection."""
        graph_ptr = np.ascontiguousarray(graph_ptr, dtype=np.int32)
        frame = np.asarray(frame)
        n = frame.shape[0]
        order = np.arange(n, dtype=np.int64)
        if sort and n:
            gid = np.repeat(np.arange(len(graph_ptr) - 1), np.diff(graph_ptr))
            order = np.lexsort((np.arange(n), frame, gid))  # stable: graph, then frame, then input position
        pos = np.asarray(pos, dtype=np.float64).reshape(n, -1)
        conf = np.full(n, 0.5) if conf is None else np.asarray(conf, dtype=np.float64)
        b = DetectionBatch(
            graph_ptr,
            np.ascontiguousarray(frame[order], dtype=np.int32),
            np.ascontiguousarray(pos[order].T),
            np.ascontiguousarray(conf[order]),
            None if flags is None else np.ascontiguousarray(np.asarray(flags, dtype=np.uint8)[order]),
            None if bbox is None else np.ascontiguousarray(np.asarray(bbox, dtype=np.float64).reshape(n, 4)[order].T),
            None if hist is None else np.ascontiguousarray(np.asarray(hist, dtype=np.float64)[order]))
        return b, order

    def max_graph_nodes(self):
        gp = self.graph_ptr.cpu().numpy() if torch.is_tensor(self.graph_ptr) else self.graph_ptr
        return int(np.diff(gp).max()) if len(gp) > 1 else 0

    def to(self, device):
        def mv(a):
            if a is None:
                return None
            t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
            return t.to(device, non_blocking=True).contiguous()
        views = None if self.views is None else {k: mv(v) for k, v in self.views.items()}
        b = DetectionBatch(mv(self.graph_ptr), mv(self.frame), mv(self.pos), mv(self.conf), mv(self.flags),
                           mv(self.bbox), mv(self.hist), views)
        b._max_nodes = self.max_graph_nodes()
        return b


class GraphBatch(object):
    """Device-resident graphs built from a DetectionBatch: time keys, per-detection arc costs and the CSR of the
    transition arcs (rows = heads when direction == DIR_IN)."""

    def __init__(self):
        self.dets = None
        self.spec = None
        self.cspec = None
        self.tkey = None
        self.entry_cost = self.node_cost = self.exit_cost = None
        self.direction = _lib.DIR_IN
        self.row_ptr = self.col = self.cost = self.weight = None
        self.n_edges = 0


def _views_mode(spec):
    """3-D detections with appearance terms: averaged over the shared 2-D views by mot3d_edge_appearance_views."""
    return getattr(spec, "kind", None) == "detections_3d" and (spec.use_hist or spec.use_bbox)


def _dets_struct(b, tkey, spec, sorted_pos=None, sorted_perm=None):
    d = _lib.Dets()
    if sorted_pos is not None:
        d.sorted_pos = sorted_pos.data_ptr()
        d.sorted_perm = sorted_perm.data_ptr()
    d.n = b.n
    d.dim = b.dim
    d.tkey = tkey.data_ptr()
    d.pos = b.pos.data_ptr()
    d.conf = b.conf.data_ptr()
    d.flags = b.flags.data_ptr() if b.flags is not None else None
    if _views_mode(spec):
        if b.views is None:
            raise ValueError("use_color_histogram / use_bbox on 3-D detections need their detections_2d views")
        return d
    d.bbox = b.bbox.data_ptr() if (b.bbox is not None and spec.use_bbox) else None
    if spec.use_hist:
        if b.hist is None:
            raise ValueError("use_color_histogram=True but the detections carry no color_histogram")
        d.hist = b.hist.data_ptr()
        d.hist_channels = int(b.hist.shape[1])
        d.hist_bins = int(b.hist.shape[2])
    if spec.use_bbox and b.bbox is None:
        raise ValueError("use_bbox=True but the detections carry no bbox")
    return d


def _raise_key_errors(flag):
    """MOT3D_ERR_UNSORTED / MOT3D_ERR_KEY_RANGE left by mot3d_make_keys (include/mot3d_b200.h)."""
    if flag & _lib.ERR_UNSORTED:
        raise ValueError("the detections of a graph are not sorted by frame (DetectionBatch.from_arrays(sort=True) "
                         "sorts them; mot3d.build_graph always does)")
    if flag & _lib.ERR_KEY_RANGE:
        raise _lib.Mot3dError(_lib.E_UNSUPPORTED, "the frame spans of the batch add up to more than 2^31 time keys: "
                              "split the batch")


def build_graphs(batch, spec, direction=_lib.DIR_IN, with_weights=False, keys=None, pruned=True, fused=True):
    """Graph build on the device. `batch` must already live on a CUDA device (DetectionBatch.to).
    pruned=True: per-frame x-sorted copies + window search; fused=True (needs pruned): the single-pass kernel
    (csrc/edge_fused.cu), else count -> scan -> fill (csrc/edge_pruned.cu). pruned=False: test every pair
    (csrc/edge_build.cu). All three produce the same arrays bit for bit."""
    L = _lib.lib()
    dev = batch.frame.device
    if dev.type != "cuda":
        raise RuntimeError("mot3d_b200.build_graphs needs CUDA tensors; there is no CPU path")
    n, G = batch.n, batch.n_graphs
    st = _stream(dev)
    cs = spec.to_c()
    cs_real = cs
    views_mode = _views_mode(spec)
    if views_mode:
        # the build leaves the distance factor in weight[]; mot3d_edge_appearance_views finishes the arcs
        cs = spec.to_c()
        cs.use_hist, cs.use_bbox, cs.defer_cost = 0, 0, 1
    gb = GraphBatch()
    gb.dets, gb.spec, gb.cspec, gb.direction = batch, spec, cs_real, direction
    i32 = dict(dtype=torch.int32, device=dev)
    i64 = dict(dtype=torch.int64, device=dev)
    overflow = torch.zeros(1, **i32)  # MOT3D_ERR_* bits of the keys and the arc build, read back once
    if keys is None:
        tkey = torch.empty(max(n, 1), **i32)
        base = torch.empty(max(G, 1), **i64)
        _lib.check(L.mot3d_make_keys(_ptr(batch.frame), _ptr(batch.graph_ptr), G, spec.max_jump, _ptr(tkey),
                                     _ptr(base), _ptr(overflow), st))
    else:
        tkey = keys
    gb.tkey = tkey
    spos = sperm = None
    if pruned and n:
        spos = torch.empty(batch.dim * n, dtype=torch.float64, device=dev)
        sperm = torch.empty(n, **i32)
    d = _dets_struct(batch, tkey, spec, spos, sperm)
    if spos is not None:
        _lib.check(L.mot3d_frame_sort(ctypes.byref(d), st))
    gb.entry_cost = torch.empty(max(n, 1), **i64)
    gb.node_cost = torch.empty(max(n, 1), **i64)
    gb.exit_cost = torch.empty(max(n, 1), **i64)
    _lib.check(L.mot3d_node_costs(ctypes.byref(d), ctypes.byref(cs), _ptr(gb.entry_cost), _ptr(gb.node_cost),
                                  _ptr(gb.exit_cost), st))
    need_w = with_weights or spec.use_hist or views_mode
    if fused and pruned and n:
        # one launch; the arc arrays are sized by a guess and the launch is repeated in the rare case it was too small
        row_ptr = torch.empty(n + 1, **i32)
        ws_bytes = L.mot3d_edge_build_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        cap = max(16 * n, 1024)
        while True:
            gb.col = torch.empty(cap, **i32)
            gb.cost = torch.empty(cap, **i64)
            gb.weight = torch.empty(cap, dtype=torch.float64, device=dev) if need_w else None
            _lib.check(L.mot3d_edge_build(ctypes.byref(d), ctypes.byref(cs), direction, 0, n, _ptr(row_ptr), cap,
                                          _ptr(gb.col), _ptr(gb.cost), _ptr(gb.weight), _ptr(overflow), None, None,
                                          _ptr(ws), ws_bytes, st))
            n_edges = int(row_ptr[n].item())  # the one host read-back of the build
            flag = int(overflow.item())
            _raise_key_errors(flag)
            if flag & _lib.ERR_WINDOW:
                raise _lib.Mot3dError(_lib.E_UNSUPPORTED, "more than 2^24 detections inside one time window")
            if n_edges <= cap:
                break
            cap = n_edges
            overflow.zero_()
        gb.n_edges = n_edges
        gb.row_ptr = row_ptr
        gb.col = gb.col[:max(n_edges, 1)]
        gb.cost = gb.cost[:max(n_edges, 1)]
        if need_w:
            gb.weight = gb.weight[:max(n_edges, 1)]
    else:
        row_count = torch.empty(max(n, 1), **i32)
        _lib.check(L.mot3d_edge_count(ctypes.byref(d), ctypes.byref(cs), direction, _ptr(row_count), st))
        row_ptr = torch.empty(n + 1, **i32)
        ws_bytes = L.mot3d_scan_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(L.mot3d_exclusive_scan_i32(_ptr(row_count), n, _ptr(row_ptr), _ptr(ws), ws_bytes, st))
        n_edges = int(row_ptr[n].item()) if n else 0  # the one host read-back of the build: sizes the arc arrays
        gb.n_edges = n_edges
        gb.row_ptr = row_ptr
        gb.col = torch.empty(max(n_edges, 1), **i32)
        gb.cost = torch.empty(max(n_edges, 1), **i64)
        gb.weight = torch.empty(max(n_edges, 1), dtype=torch.float64, device=dev) if need_w else None
        _raise_key_errors(int(overflow.item()))
        _lib.check(L.mot3d_edge_fill(ctypes.byref(d), ctypes.byref(cs), direction, _ptr(row_ptr), n_edges,
                                     _ptr(gb.col), _ptr(gb.cost), _ptr(gb.weight), _ptr(overflow), st))
    if views_mode:
        if n:
            v = batch.views
            vs = _lib.Views()
            vs.n, vs.n_entries = n, int(v["view_id"].shape[0])
            vs.view_ptr, vs.view_id = v["view_ptr"].data_ptr(), v["view_id"].data_ptr()
            if spec.use_bbox:
                vs.bbox = v["bbox"].data_ptr()
            if spec.use_hist:
                vs.hist = v["hist"].data_ptr()
                vs.hist_channels, vs.hist_bins = int(v["hist"].shape[1]), int(v["hist"].shape[2])
            err = torch.zeros(1, **i32)
            _lib.check(L.mot3d_edge_appearance_views(ctypes.byref(d), ctypes.byref(vs), ctypes.byref(cs_real), direction,
                                                     _ptr(gb.row_ptr), _ptr(gb.col), n_edges, _ptr(gb.weight),
                                                     _ptr(gb.cost), _ptr(err), st))
            if int(err.item()):
                raise ValueError("two linked 3-D detections share no 2-D view: weight_distance_detections_3d is nan "
                                 "for them (np.mean of an empty list, reference mot3d/weight_functions.py:55-71)")
    elif spec.use_hist:
        _lib.check(L.mot3d_edge_appearance(ctypes.byref(d), ctypes.byref(cs), direction, _ptr(row_ptr), _ptr(gb.col),
                                           n_edges, _ptr(gb.weight), _ptr(gb.cost), st))
    return gb


class Solution(object):
    """next/prev per detection (global indices; -2 terminal, -1 unused), K and total cost per graph."""

    def __init__(self):
        self.next = self.prev = self.n_paths = self.total_cost = self.path_cost = None
        self.track_count = self.track_ptr = self.track_det = None


def solve_graphs(gb, want_path_costs=False):
    L = _lib.lib()
    if gb.direction != _lib.DIR_IN:
        raise ValueError("the solver pulls over in-arcs: build with direction=DIR_IN")
    b = gb.dets
    dev = b.frame.device
    n, G = b.n, b.n_graphs
    st = _stream(dev)
    i32 = dict(dtype=torch.int32, device=dev)
    i64 = dict(dtype=torch.int64, device=dev)
    sol = Solution()
    sol.next = torch.empty(max(n, 1), **i32)
    sol.prev = torch.empty(max(n, 1), **i32)
    sol.n_paths = torch.zeros(max(G, 1), **i32)
    sol.total_cost = torch.zeros(max(G, 1), **i64)
    sol.path_cost = torch.zeros(max(n, 1), **i64) if want_path_costs else None
    sol.stats = torch.zeros(max(G, 1) * _lib.SOLVE_STATS, **i64)
    ws_bytes = L.mot3d_solve_workspace_bytes(n, G, max(gb.n_edges, 0))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    max_nodes = getattr(b, "_max_nodes", None)
    if max_nodes is None:
        max_nodes = b.max_graph_nodes()
    _lib.check(L.mot3d_solve_batch(G, n, _ptr(b.graph_ptr), max_nodes, _ptr(gb.tkey), max(gb.n_edges, 0),
                                   _ptr(gb.row_ptr), _ptr(gb.col),
                                   _ptr(gb.cost), _ptr(gb.entry_cost), _ptr(gb.node_cost), _ptr(gb.exit_cost),
                                   _ptr(sol.next), _ptr(sol.prev), _ptr(sol.n_paths), _ptr(sol.total_cost),
                                   _ptr(sol.path_cost), _ptr(sol.stats), _ptr(ws), ws_bytes, st))
    if want_path_costs and n:
        # the kernel leaves the accepted path costs grouped by connected component inside every graph's slice
        # (MOT3D_INF_COST elsewhere); present them per graph in the order the reference finds them (ascending)
        scratch = torch.empty(n, **i64)
        many = torch.zeros(1, **i32)
        _lib.check(L.mot3d_order_path_costs(G, _ptr(b.graph_ptr), _ptr(sol.path_cost), _ptr(scratch), 8192, _ptr(many),
                                            st))
        if int(many.item()):  # a graph with more than 8192 tracks: not worth a kernel of its own
            gp = b.graph_ptr.to(torch.int64)
            gid = torch.repeat_interleave(torch.arange(G, device=dev), gp[1:] - gp[:-1])
            v, o = torch.sort(sol.path_cost[:n], stable=True)
            _, o2 = torch.sort(gid[o], stable=True)
            sol.path_cost = torch.cat([v[o2], sol.path_cost[n:]])
    return sol


def extract_tracks(gb, sol):
    L = _lib.lib()
    b = gb.dets
    dev = b.frame.device
    n, G = b.n, b.n_graphs
    i32 = dict(dtype=torch.int32, device=dev)
    sol.track_count = torch.zeros(max(G, 1), **i32)
    sol.track_ptr = torch.zeros(n + G + 1, **i32)
    sol.track_det = torch.zeros(max(n, 1), **i32)
    _lib.check(L.mot3d_extract_tracks(G, _ptr(b.graph_ptr), _ptr(sol.next), _ptr(sol.prev), _ptr(sol.track_count),
                                      _ptr(sol.track_ptr), _ptr(sol.track_det), _stream(dev)))
    return sol


def tracks_to_lists(graph_ptr, track_count, track_ptr, track_det):
    """Host arrays -> per graph list of tracks, each a list of detection indices local to the graph."""
    graph_ptr = np.asarray(graph_ptr)
    out = []
    for g in range(len(graph_ptr) - 1):
        g0 = int(graph_ptr[g])
        K = int(track_count[g])
        tp = track_ptr[g0 + g: g0 + g + K + 1]
        td = track_det[g0:]
        out.append([td[int(tp[k]):int(tp[k + 1])].tolist() for k in range(K)])
    return out


def track_batch(batch, spec, want_path_costs=False):
    """build + solve + extract on the device; returns (GraphBatch, Solution)."""
    gb = build_graphs(batch, spec)
    sol = solve_graphs(gb, want_path_costs=want_path_costs)
    extract_tracks(gb, sol)
    return gb, sol


class HostTracker(object):
    """Reusable pinned host buffers + device workspace for mot3d_track_batch_host (one C-ABI call per batch, host
    buffers in, host buffers out, copies inside)."""

    def __init__(self, n_max, g_max, edge_capacity, dim=2, device=None, with_bbox=False):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_max, self.g_max, self.edge_capacity, self.dim = int(n_max), int(g_max), int(edge_capacity), dim
        L = _lib.lib()
        self.ws_bytes = L.mot3d_track_batch_workspace_bytes(self.n_max, self.g_max, self.edge_capacity)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        pin = dict(pin_memory=True)
        self.h_graph_ptr = torch.zeros(g_max + 1, dtype=torch.int32, **pin)
        self.h_frame = torch.zeros(n_max, dtype=torch.int32, **pin)
        self.h_pos = torch.zeros(dim * n_max, dtype=torch.float64, **pin)
        self.h_conf = torch.zeros(n_max, dtype=torch.float64, **pin)
        self.h_bbox = torch.zeros(4 * n_max, dtype=torch.float64, **pin) if with_bbox else None
        self.h_track_count = torch.zeros(g_max, dtype=torch.int32, **pin)
        self.h_track_ptr = torch.zeros(n_max + g_max, dtype=torch.int32, **pin)
        self.h_track_det = torch.zeros(n_max, dtype=torch.int32, **pin)
        self.h_total = torch.zeros(g_max, dtype=torch.int64, **pin)

    def load(self, batch):
        """Copy a host DetectionBatch (numpy) into the pinned staging buffers (not part of the C-ABI call)."""
        n, G = batch.n, batch.n_graphs
        assert n <= self.n_max and G <= self.g_max and batch.dim == self.dim
        self.n, self.G = n, G
        self.h_graph_ptr[:G + 1] = torch.from_numpy(np.asarray(batch.graph_ptr))
        self.h_frame[:n] = torch.from_numpy(np.asarray(batch.frame))
        self.h_pos[:self.dim * n] = torch.from_numpy(np.asarray(batch.pos).reshape(-1))
        self.h_conf[:n] = torch.from_numpy(np.asarray(batch.conf))
        if self.h_bbox is not None and batch.bbox is not None:
            self.h_bbox[:4 * n] = torch.from_numpy(np.asarray(batch.bbox).reshape(-1))

    def h2d_bytes(self):
        """Bytes the last run() copied host -> device (counted by the C call)."""
        return self._h2d

    def d2h_bytes(self):
        """Bytes the last run() copied device -> host."""
        return self._d2h

    def run(self, spec_c):
        """One mot3d_track_batch_host call on the staged batch. Returns the number of transition arcs."""
        L = _lib.lib()
        hb = _lib.HostBatch()
        hb.n_graphs, hb.dim, hb.n = self.G, self.dim, self.n
        hb.graph_ptr = self.h_graph_ptr.data_ptr()
        hb.frame = self.h_frame.data_ptr()
        hb.pos = self.h_pos.data_ptr()
        hb.conf = self.h_conf.data_ptr()
        hb.flags = None
        hb.bbox = self.h_bbox.data_ptr() if self.h_bbox is not None else None
        res = _lib.HostResult()
        res.track_count = self.h_track_count.data_ptr()
        res.track_ptr = self.h_track_ptr.data_ptr()
        res.track_det = self.h_track_det.data_ptr()
        res.total_cost = self.h_total.data_ptr()
        with torch.cuda.device(self.device):
            _lib.check(L.mot3d_track_batch_host(ctypes.byref(hb), ctypes.byref(spec_c), ctypes.byref(res),
                                                _ptr(self.ws), self.ws_bytes, self.edge_capacity,
                                                _stream(self.device)))
        self._h2d, self._d2h = int(res.h2d_bytes), int(res.d2h_bytes)
        return int(res.n_edges)

    def close(self):
        """Release the streams / events / pinned scratch mot3d_track_batch_host keeps for this host thread
        (mot3d_host_release); the next run() creates them again."""
        torch.cuda.synchronize(self.device)
        _lib.check(_lib.lib().mot3d_host_release())

    def tracks(self):
        return tracks_to_lists(self.h_graph_ptr[:self.G + 1].numpy(), self.h_track_count.numpy(),
                               self.h_track_ptr.numpy(), self.h_track_det.numpy())
