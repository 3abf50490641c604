"""The reference's public API (reference mot3d/mot3d.py:18: build_graph, solve_graph, graph_to_text, save_graph),
backed by the CUDA kernels. Same names, arguments, defaults, return conventions and error behaviour:

    g = build_graph(detections, weight_source_sink=1, max_jump=12, verbose=True,
                    weight_confidence=..., weight_distance=...)        -> graph handle or None
    trajectories = solve_graph(g, verbose=True, method='muSSP')         -> list[list[Detection]]

The graph handle owns device tensors; it materialises a networkx.DiGraph identical to the reference's only when a
caller asks for nodes/edges (plot_graph, graph_to_text).
"""
import os
import time

import numpy as np

from . import _lib
from . import spec as _spec
from .types import Detection, Detection2D, Detection3D, DetectionTracklet
from .weight_functions import weight_distance_detections_2d, weight_confidence_detections_2d

__all__ = ["build_graph", "solve_graph", "graph_to_text", "save_graph"]


class _ProbeMixin(object):
    _m3_probe = True


class _Probe2D(_ProbeMixin, Detection2D):
    pass


class _Probe3D(_ProbeMixin, Detection3D):
    pass


def _recover_spec(detections, weight_source_sink, max_jump, weight_confidence, weight_distance):
    """Find out which built-in costs (and kwargs) the user's callables stand for; see spec.probe_callable."""
    first = detections[0]
    if isinstance(first, Detection3D):
        p1, p2 = _Probe3D(0, (0.0, 0.0, 0.0)), _Probe3D(1, (0.0, 0.0, 0.0))
    else:
        p1, p2 = _Probe2D(0, (0.0, 0.0)), _Probe2D(1, (0.0, 0.0))
    rd = _spec.probe_callable(weight_distance, p1, p2)
    rc = _spec.probe_callable(weight_confidence, p1)
    if rc.name != "confidence":
        raise TypeError("weight_confidence must be one of the built-in weight_confidence_* functions")
    if rd.name not in ("detections_2d", "detections_3d"):
        raise TypeError("weight_distance resolved to %r, which does not apply to detections" % rd.name)
    return _spec.CostSpec(kind=rd.name, distance=rd.kwargs, confidence=rc.kwargs,
                          weight_source_sink=weight_source_sink, max_jump=max_jump)


def _recover_tracklet_spec(tracklets, weight_source_sink, max_jump, weight_confidence, weight_distance):
    from .types import DetectionTracklet2D, DetectionTracklet3D

    class _PT2(_ProbeMixin, DetectionTracklet2D):
        def __init__(self):
            pass

    class _PT3(_ProbeMixin, DetectionTracklet3D):
        def __init__(self):
            pass

    cls = _PT3 if isinstance(tracklets[0], DetectionTracklet3D) else _PT2
    rd = _spec.probe_callable(weight_distance, cls(), cls())
    rc = _spec.probe_callable(weight_confidence, cls())
    if rc.name != "confidence":
        raise TypeError("weight_confidence must be one of the built-in weight_confidence_* functions")
    if rd.name not in ("tracklets_2d", "tracklets_3d"):
        raise TypeError("weight_distance resolved to %r, which does not apply to tracklets" % rd.name)
    return _spec.TrackletCostSpec(kind=rd.name, distance=rd.kwargs, confidence=rc.kwargs,
                                  weight_source_sink=weight_source_sink, max_jump=max_jump)


class TrackletGraph(object):
    """Graph handle for a list of DetectionTracklet objects (nodes in the caller's order, reference
    mot3d/mot3d.py:134-135). Same surface as TrackingGraph."""

    def __init__(self, tracklets, tspec, packed, out):
        self.detections = tracklets
        self.spec = tspec
        self.packed = packed
        self._out = out  # (row_ptr, col, weight, cost) on the host, reference arc order
        n = len(tracklets)
        flags = packed["flags"]
        en = np.full(n, _spec.quantise_scalar(tspec.weight_source_sink), dtype=np.int64)
        ex = np.zeros(n, dtype=np.int64)
        if flags is not None:
            en[(flags & _lib.FLAG_ENTRY) == 0] = _lib.INF_COST
            ex[(flags & _lib.FLAG_EXIT) == 0] = _lib.INF_COST
        self.entry_cost, self.exit_cost = en, ex
        self.node_weight = -(packed["conf"] * tspec.confidence["mul"] + tspec.confidence["bias"])
        self.node_cost = np.array([_spec.quantise_scalar(w) for w in self.node_weight.tolist()], dtype=np.int64)
        self._nx = None
        self._solution = None

    def __len__(self):
        return 2 * len(self.detections) + 2

    @property
    def n_transitions(self):
        return len(self._out[1])

    def out_csr(self):
        return self._out

    def arcs(self):
        n = len(self.detections)
        row_ptr, col, w, _ = self._out
        return _arcs_in_reference_order(n, row_ptr, col, w, self.entry_cost < _lib.INF_COST,
                                        self.exit_cost < _lib.INF_COST, self.spec.weight_source_sink, self.node_weight)

    def solve(self):
        from .tracklets import solve_tracklet_graph
        row_ptr, col, _, cost = self._out
        tracks, total, K, sol = solve_tracklet_graph(len(self.detections), row_ptr, col, cost, self.entry_cost,
                                                     self.node_cost, self.exit_cost)
        self._solution = sol
        self.total_cost, self.n_paths = total, K
        return tracks

    to_networkx = None  # bound below (shared with TrackingGraph)


def _arcs_in_reference_order(n, row_ptr, col, w, has_en, has_ex, weight_source_sink, node_w):
    """Every arc as (tail, head, weight), 1-based node ids, in graph_to_text order (reference
    mot3d/mot3d.py:75-104,218-226): all S->pre first, then per item pre->post, post->T, post->pre_j (j ascending)."""
    T = 2 * n + 2
    m = np.arange(n)
    deg = np.diff(row_ptr)
    per = 1 + has_ex.astype(np.int64) + deg
    n_en = int(has_en.sum())
    E = n_en + int(per.sum())
    tail = np.empty(E, dtype=np.int64)
    head = np.empty(E, dtype=np.int64)
    ww = np.empty(E, dtype=np.float64)
    tail[:n_en] = 1
    head[:n_en] = 2 + 2 * m[has_en]
    ww[:n_en] = weight_source_sink
    base = n_en + np.concatenate(([0], np.cumsum(per)[:-1])) if n else np.zeros(0, dtype=np.int64)
    tail[base] = 2 + 2 * m
    head[base] = 3 + 2 * m
    ww[base] = node_w
    xi = base[has_ex] + 1
    tail[xi] = 3 + 2 * m[has_ex]
    head[xi] = T
    ww[xi] = 0
    if len(col):
        src = np.repeat(m, deg)
        within = np.arange(len(col)) - np.repeat(row_ptr[:-1], deg)
        at = np.repeat(base + 1 + has_ex.astype(np.int64), deg) + within
        tail[at] = 3 + 2 * src
        head[at] = 2 + 2 * col
        ww[at] = w
    return tail, head, ww


def _pack(sorted_dets, cspec):
    """Detection objects (already in node order) -> DetectionBatch on the host (one graph)."""
    from .batch import DetectionBatch
    n = len(sorted_dets)
    dim = cspec.dim
    idx = [d.index for d in sorted_dets]
    frame = np.fromiter(idx, dtype=np.int64, count=n)
    if any(f != i for f, i in zip(frame.tolist(), idx)):
        raise ValueError("detection.index must be an integer frame number (the time keys of the arc kernels are "
                         "integers); got a fractional index")
    if n and (frame.min() < -2 ** 31 or frame.max() >= 2 ** 31):
        raise ValueError("frame indices must fit in int32")
    try:  # one conversion for the usual case: every position is a flat sequence of `dim` numbers
        pos = np.array([d.position for d in sorted_dets], dtype=np.float64)
    except ValueError:
        pos = None
    if pos is None or pos.ndim != 2 or pos.shape[1] != dim:
        pos = np.empty((n, dim), dtype=np.float64)
        for i, d in enumerate(sorted_dets):
            q = np.asarray(d.position, dtype=np.float64).reshape(-1)
            if q.shape[0] != dim:
                # the reference's euclidean() takes whatever coordinates there are; the kernels take exactly `dim`
                raise ValueError("detection %d has a %d-D position but the cost function is for %d-D detections" %
                                 (i, q.shape[0], dim))
            pos[i] = q
    conf = np.fromiter((d.confidence for d in sorted_dets), dtype=np.float64, count=n)
    flags = None
    if any((not d.entry_points) or (not d.exit_point) for d in sorted_dets):
        flags = np.fromiter(((_lib.FLAG_ENTRY if d.entry_points else 0) | (_lib.FLAG_EXIT if d.exit_point else 0)
                             for d in sorted_dets), dtype=np.uint8, count=n)
    bbox = hist = views = None
    if cspec.kind == "detections_3d" and (cspec.use_bbox or cspec.use_hist):
        views = _pack_views(sorted_dets, cspec)
    else:
        if cspec.use_bbox:
            bbox = np.array([d.bbox for d in sorted_dets], dtype=np.float64).reshape(n, 4)
        if cspec.use_hist:
            hist = np.array([np.asarray(d.color_histogram, dtype=np.float64) for d in sorted_dets])
            if hist.ndim != 3:
                raise ValueError("color_histogram must be a list of C arrays of B bins for every detection")
    gp = np.array([0, n], dtype=np.int32)
    return DetectionBatch(gp, frame.astype(np.int32), np.ascontiguousarray(pos.T), conf, flags,
                          None if bbox is None else np.ascontiguousarray(bbox.T), hist, views)


def _pack_views(sorted_dets, cspec):
    """Detection3D.detections_2d ({view: Detection2D}, reference mot3d/types.py:40-45) -> packed entries in each
    detection's dict order, which is the order weight_distance_detections_3d averages in
    (reference mot3d/weight_functions.py:55-71). View labels are mapped to integers."""
    labels = {}
    ptr, ids, bbs, hs = [0], [], [], []
    for d in sorted_dets:
        d2 = getattr(d, "detections_2d", None) or {}
        for view, det in d2.items():
            ids.append(labels.setdefault(view, len(labels)))
            if cspec.use_bbox:
                bbs.append(np.asarray(det.bbox, dtype=np.float64).reshape(4))
            if cspec.use_hist:
                hs.append(np.asarray(det.color_histogram, dtype=np.float64))
        ptr.append(len(ids))
    views = dict(view_ptr=np.array(ptr, dtype=np.int32), view_id=np.array(ids, dtype=np.int32))
    if cspec.use_bbox:
        views["bbox"] = np.ascontiguousarray(np.array(bbs, dtype=np.float64).reshape(len(ids), 4).T)
    if cspec.use_hist:
        h = np.array(hs, dtype=np.float64)
        if len(ids) and h.ndim != 3:
            raise ValueError("color_histogram must be a list of C arrays of B bins for every 2-D view")
        views["hist"] = np.ascontiguousarray(h.reshape(len(ids), *h.shape[1:]) if len(ids) else np.zeros((0, 1, 1)))
    return views


class TrackingGraph(object):
    """What build_graph returns. Quacks like the reference's networkx graph where callers touch it:
    len(g) == number of nodes; g.edges(), g.nodes[...], g.has_edge(...) etc. are served by a lazily built
    networkx.DiGraph with the reference's node ids, attributes and edge order."""

    def __init__(self, detections, cspec, graphs):
        self.detections = detections  # the caller's objects, in node order
        self.spec = cspec
        self.graphs = graphs          # batch.GraphBatch (one graph, DIR_IN CSR on the device)
        self._nx = None
        self._out = None
        self._solution = None

    def __len__(self):
        return 2 * len(self.detections) + 2

    @property
    def n_transitions(self):
        return self.graphs.n_edges

    def out_csr(self):
        """(row_ptr, col, weight float64, cost int64) on the host, rows = tails, reference emission order."""
        if self._out is None:
            from .batch import build_graphs
            gb = build_graphs(self.graphs.dets, self.spec, direction=_lib.DIR_OUT, with_weights=True,
                              keys=self.graphs.tkey)
            e = gb.n_edges
            self._out = (gb.row_ptr.cpu().numpy().astype(np.int64), gb.col[:e].cpu().numpy().astype(np.int64),
                         gb.weight[:e].cpu().numpy(), gb.cost[:e].cpu().numpy())
        return self._out

    def arcs(self):
        """Every arc as (tail, head, weight) arrays, 1-based node ids, in graph_to_text order
        (reference mot3d/mot3d.py:75-104,218-226)."""
        n = len(self.detections)
        row_ptr, col, w, _ = self.out_csr()
        gb = self.graphs
        en = gb.entry_cost[:n].cpu().numpy()
        ex = gb.exit_cost[:n].cpu().numpy()
        conf = gb.dets.conf.cpu().numpy()
        node_w = -(conf * self.spec.confidence["mul"] + self.spec.confidence["bias"])
        return _arcs_in_reference_order(n, row_ptr, col, w, en < _lib.INF_COST, ex < _lib.INF_COST,
                                        self.spec.weight_source_sink, node_w)

    def solve(self):
        """Tracks as lists of detection indices (node order), in the reference's order."""
        from .batch import solve_graphs, extract_tracks, tracks_to_lists
        sol = self._solution if self._solution is not None else solve_graphs(self.graphs)
        extract_tracks(self.graphs, sol)
        self._solution = sol
        n = len(self.detections)
        return tracks_to_lists(np.array([0, n]), sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(),
                               sol.track_det.cpu().numpy())[0]

    def to_networkx(self):
        """The reference's DiGraph: nodes SOURCE=1, SINK=2N+2, pre=2+2m, post=3+2m with label/detection/idet
        attributes, edges with 'weight', same insertion order (reference mot3d/mot3d.py:75-141)."""
        if self._nx is None:
            import networkx as nx
            n = len(self.detections)
            g = nx.DiGraph()
            g.add_node(1, label="source")
            g.add_node(2 * n + 2, label="sink")
            for m, d in enumerate(self.detections):
                g.add_node(2 + 2 * m, detection=d, label="pre-node", idet=m)
                g.add_node(3 + 2 * m, detection=d, label="post-node", idet=m)
            tail, head, w = self.arcs()
            g.add_weighted_edges_from(zip(tail.tolist(), head.tolist(), w.tolist()))
            self._nx = g
        return self._nx

    def __getattr__(self, name):  # nodes, edges, has_edge, ... -> the materialised networkx graph
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.to_networkx(), name)

    def __iter__(self):
        return iter(self.to_networkx())

    def __contains__(self, node):
        return 1 <= node <= len(self)

    def __getitem__(self, node):
        return self.to_networkx()[node]


for _name in ("to_networkx", "__getattr__", "__iter__", "__contains__", "__getitem__"):
    setattr(TrackletGraph, _name, TrackingGraph.__dict__[_name])


def _build_tracklet_graph(tracklets, weight_source_sink, max_jump, verbose, weight_confidence, weight_distance):
    from .tracklets import pack_tracklets, build_tracklet_arcs
    if any(type(t) is not type(tracklets[0]) for t in tracklets):
        raise NotImplementedError("build_graph on a list that mixes 2-D and 3-D tracklets: the reference's mixed costs "
                                  "cannot run either (mot3d/weight_functions.py:174)")
    tspec = _recover_tracklet_spec(tracklets, weight_source_sink, max_jump, weight_confidence, weight_distance)
    n = len(tracklets)
    if verbose:
        print("Number of tracklets: {}".format(n))
    node = 2
    for t in tracklets:
        t.pre_node = node
        t.post_node = node + 1
        node += 2
    packed = pack_tracklets(tracklets, tspec)
    out = build_tracklet_arcs(packed, tspec)
    g = TrackletGraph(list(tracklets), tspec, packed, out)
    if verbose:
        print("Number of post-pre nodes edges created: {}".format(len(out[1])))
    if packed["flags"] is not None:
        g._tracks = g.solve()
        if g.n_paths == 0:
            return None
    return g


def build_graph(detections, weight_source_sink=1, max_jump=12, verbose=True,
                weight_confidence=weight_confidence_detections_2d, weight_distance=weight_distance_detections_2d):
    """Build the flow graph on the GPU (reference mot3d/mot3d.py:20-150).

    detections : list/tuple of Detection2D / Detection3D in any order.
    Returns a TrackingGraph, or None when there is no detection or no path from source to sink, exactly like the
    reference (mot3d/mot3d.py:49-50,147-150). Sets detection.pre_node / detection.post_node (:89,:97).
    """
    assert isinstance(detections, (list, tuple))
    if len(detections) == 0:
        return None
    import torch
    if isinstance(detections[0], DetectionTracklet):
        if not torch.cuda.is_available():
            raise RuntimeError("mot3d_b200.build_graph needs a CUDA device (B200); there is no CPU fallback")
        return _build_tracklet_graph(detections, weight_source_sink, max_jump, verbose, weight_confidence,
                                     weight_distance)
    if not isinstance(detections[0], Detection):
        raise TypeError("detections must be Detection2D/Detection3D objects")
    from .batch import build_graphs, solve_graphs
    if not torch.cuda.is_available():
        raise RuntimeError("mot3d_b200.build_graph needs a CUDA device (B200); there is no CPU fallback")
    cspec = _recover_spec(detections, weight_source_sink, max_jump, weight_confidence, weight_distance)
    # stable ascending sort by frame index = the reference's cmp sort on -a.diff_index(b) (mot3d/mot3d.py:61-64)
    ordered = sorted(detections, key=lambda d: d.index)
    n = len(ordered)
    if verbose:
        print("Number of frames: {}".format(len(set(d.index for d in ordered))))
        print("Number of detections: {}".format(n))
    node = 2
    for d in ordered:
        d.pre_node = node
        d.post_node = node + 1
        node += 2
    host = _pack(ordered, cspec)
    dev = torch.device("cuda", torch.cuda.current_device())
    graphs = build_graphs(host.to(dev), cspec)
    g = TrackingGraph(ordered, cspec, graphs)
    if verbose:
        print("Number of post-pre nodes edges created: {}".format(graphs.n_edges))
        n_arcs = graphs.n_edges + 3 * n if host.flags is None else None
        if n_arcs is not None:
            print("Fraction of edges per detection: {}".format(n_arcs / n))
    if host.flags is not None:
        # a source->sink path exists iff the solver finds a first path; keep the solution for solve_graph
        sol = solve_graphs(graphs)
        if int(sol.n_paths[0].item()) == 0:
            return None
        g._solution = sol
    return g


def _run_ssp(g, verbose=1, method="muSSP"):
    """GPU successive shortest paths + flow following -> list of tracks as lists of node ids (post-nodes), in the
    reference's order (reference mot3d/mot3d.py:152-184)."""
    if method != "muSSP":
        raise NotImplementedError(method)
    tracks = getattr(g, "_tracks", None)
    if tracks is None:
        tracks = g.solve()
    n = len(g.detections)
    if verbose:
        used = sum(len(t) for t in tracks) * 2
        print("[muSSP] {} out of {} nodes unused".format(2 * n - used, 2 * n))
    return [[3 + 2 * m for m in t] for t in tracks]


def solve_graph(g, verbose=True, method="muSSP"):
    """Solve the flow graph (reference mot3d/mot3d.py:197-216). method 'muSSP' runs the CUDA solver;
    'FollowMe'/'SSP' raise NotImplementedError as in the reference (:159-163); 'ILP' is the reference's slow
    PuLP path and is not provided; anything else raises ValueError (:205-206)."""
    start_time = time.time()
    if method in ["muSSP", "FollowMe", "SSP"]:
        if not isinstance(g, (TrackingGraph, TrackletGraph)):
            raise TypeError("solve_graph expects the graph returned by mot3d_b200.build_graph")
        tracks_nodes = _run_ssp(g, verbose, method)
    elif method == "ILP":
        raise NotImplementedError("method='ILP' (reference mot3d/solvers/ilp/ilp.py, PuLP/CBC) is outside the "
                                  "GPU hot path; use method='muSSP'")
    else:
        raise ValueError("Unrecognazed method '{}'. Choose 'muSSP' or 'ILP'".format(method))
    trajectories = [[g.detections[(node - 3) // 2] for node in nodes] for nodes in tracks_nodes]
    if verbose:
        print("Graph solved in {:0.4f}s using {}.".format(time.time() - start_time, method))
    return trajectories


def graph_to_text(g):
    """DIMACS-like lines, weights with 7 decimals, arcs in the graph's edge order (reference mot3d/mot3d.py:218-226)."""
    if isinstance(g, (TrackingGraph, TrackletGraph)):
        tail, head, w = g.arcs()
        lines = ["p min {} {}".format(len(g), len(tail))]
        for s, t, x in zip(tail.tolist(), head.tolist(), w.tolist()):
            lines.append("a {} {} {:0.7f}".format(s, t, x))
        return lines
    lines = ["p min {} {}".format(len(g), len(g.edges()))]
    for s, t, data in g.edges(data=True):
        lines.append("a {} {} {:0.7f}".format(s, t, data["weight"]))
    return lines


def save_graph(graph_as_text, filename="/tmp/graph.txt"):
    """Write the lines to a file (reference mot3d/mot3d.py:255-265)."""
    flags = os.O_RDWR | os.O_CREAT | os.O_TRUNC
    if os.name != "nt":
        flags |= os.O_SYNC
    fd = os.open(filename, flags)
    try:
        for line in graph_as_text:
            os.write(fd, str.encode(line + "\n"))
    finally:
        os.close(fd)


def _evaluate_pair(kind, d1, d2, kwargs):
    """weight_distance_detections_*(d1, d2, **kwargs) for one real pair, on the GPU: a 2-detection graph."""
    import torch
    from .batch import build_graphs
    jump = int(d2.index - d1.index)
    if not (1 <= jump <= _lib.MAX_JUMP):
        raise ValueError("weight of a pair is only defined here for 1 <= jump <= %d (got %d)" % (_lib.MAX_JUMP, jump))
    cs = _spec.CostSpec(kind=kind, distance=kwargs, max_jump=jump)
    host = _pack([d1, d2], cs)
    gb = build_graphs(host.to(torch.device("cuda", torch.cuda.current_device())), cs, direction=_lib.DIR_OUT,
                      with_weights=True)
    if gb.n_edges == 0:
        return None
    return float(gb.weight[0].item())


def _evaluate_tracklet_pair(kind, t1, t2, kwargs):
    """weight_distance_tracklets_*(t1, t2, **kwargs) for one real pair, on the GPU: the two-tracklet arc build
    (csrc/tracklet.cu). Like the reference it needs t2.tail to start after t1.head ends
    (mot3d/distance_functions.py:49)."""
    from .tracklets import pack_tracklets, build_tracklet_arcs
    i1, i2 = t1.head.indexes, t2.tail.indexes
    assert i2[0] > i1[-1], "idxs2[0]={} idxs1[-1]={} tracklet2 cannot be more recent than tracklet1!".format(i2[0], i1[-1])
    ts = _spec.TrackletCostSpec(kind=kind, distance=kwargs, max_jump=2 ** 30)
    row_ptr, col, weight, _ = build_tracklet_arcs(pack_tracklets([t1, t2], ts), ts)
    for e in range(int(row_ptr[0]), int(row_ptr[1])):
        if int(col[e]) == 1:
            return float(weight[e])
    return None
