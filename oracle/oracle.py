"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the mot3d tracking hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module. The product (mot3d_b200/) never does; it fails loudly when its CUDA library is missing.

Contents (each function cites the reference lines it restates; /root/reference = cvlab-epfl/mot3d):
  * numpy restatement of the graph build for Detection2D/3D inputs (mot3d/mot3d.py:58-141) with the built-in
    costs (mot3d/weight_functions.py:10-78, mot3d/distance_functions.py:21-45),
  * the reference's quantisation (graph_to_text, mot3d/mot3d.py:218-226: "{:0.7f}") -> int64 units of 1e-7,
  * arc emission in graph_to_text order, so text/sha can be compared with the reference,
  * ctypes bindings to the C restatement of the solve (oracle/ssp_oracle.c) and, when it was built from the
    vendored sources (oracle/build_ref.py), to the real muSSP (oracle/_ref/libmussp_ref.so),
  * flow -> trajectories as _run_ssp does (mot3d/mot3d.py:165-184).

Parity pinning: the reference has no test suite and asserts nothing, so the oracle is
pinned to outputs of the reference itself run in the build container: tests/golden/*.npz, produced by
tests/golden/make_golden.py (imports /root/reference/mot3d unmodified + oracle/_ref muSSP).
"""
import ctypes
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(HERE, "_build")
_ORACLE_SO = os.path.join(_BUILD, "liboracle.so")
_REF_SO = os.path.join(HERE, "_ref", "libmussp_ref.so")
UNIT = 1e-7  # graph_to_text precision, mot3d/mot3d.py:224
SCALE = 10 ** 7

_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


# ----------------------------------------------------------------------------------------------- build
def build(force=False):
    """Compile oracle/ssp_oracle.c -> oracle/_build/liboracle.so (gcc, seconds)."""
    src = os.path.join(HERE, "ssp_oracle.c")
    if force or not os.path.exists(_ORACLE_SO) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-o", _ORACLE_SO, src])
    return _ORACLE_SO


_oracle_lib = None
_ref_lib = None


def _oracle():
    global _oracle_lib
    if _oracle_lib is None:
        lib = ctypes.CDLL(build())
        lib.oracle_ssp_solve.restype = ctypes.c_int
        lib.oracle_ssp_solve.argtypes = [ctypes.c_int32, ctypes.c_int32, _i32p, _i32p, _i64p, _i32p, _i32p, _i64p,
                                         ctypes.c_int32, _i64p]
        lib.oracle_ssp_solve_batched.restype = ctypes.c_int
        lib.oracle_ssp_solve_batched.argtypes = lib.oracle_ssp_solve.argtypes + [_i32p]
        _oracle_lib = lib
    return _oracle_lib


def have_ref():
    return os.path.exists(_REF_SO)


def _ref():
    global _ref_lib
    if _ref_lib is None:
        lib = ctypes.CDLL(_REF_SO)
        lib.mussp_ref_solve.restype = ctypes.c_int
        lib.mussp_ref_solve.argtypes = [ctypes.c_int, ctypes.c_int, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p,
                                        ctypes.c_int, _f64p, _f64p]
        lib.mussp_ref_solve_text.restype = ctypes.c_int
        lib.mussp_ref_solve_text.argtypes = [ctypes.c_char_p, _i32p, ctypes.c_int, _i32p, _f64p, ctypes.c_int, _f64p,
                                             _f64p, _f64p]
        _ref_lib = lib
    return _ref_lib


def _p(a, t):
    return a.ctypes.data_as(t)


# ------------------------------------------------------------------------------------------ quantisation
def quantise(w):
    """float weights -> int64 units of 1e-7, through the very text the reference prints
    (graph_to_text, mot3d/mot3d.py:224: "{:0.7f}".format(weight))."""
    w = np.asarray(w, dtype=np.float64).ravel()
    out = np.empty(w.shape[0], dtype=np.int64)
    for i, x in enumerate(w.tolist()):
        s = "{:0.7f}".format(x)
        neg = s.startswith("-")
        whole, frac = s.lstrip("-").split(".")
        v = int(whole) * SCALE + int(frac)
        out[i] = -v if neg else v
    return out


def units_to_text(c):
    """int64 units -> the "{:0.7f}" text of the reference (sign of a zero is lost; the reference may print -0.0000000)."""
    c = int(c)
    s = "-" if c < 0 else ""
    c = abs(c)
    return "%s%d.%07d" % (s, c // SCALE, c % SCALE)


# ------------------------------------------------------------------------------------------- cost model
class CostSpec(object):
    """Parameters of weight_distance_detections_2d/_3d (mot3d/weight_functions.py:13-17, 43-47),
    weight_confidence_detections (:37-38) and build_graph (mot3d/mot3d.py:20-23)."""

    def __init__(self, sigma_jump=1, sigma_distance=2, sigma_color_histogram=0.3, sigma_box_size=0.3,
                 max_distance=20, use_color_histogram=False, use_bbox=False, mul=1, bias=0,
                 weight_source_sink=1, max_jump=12):
        self.sigma_jump = sigma_jump
        self.sigma_distance = sigma_distance
        self.sigma_color_histogram = sigma_color_histogram
        self.sigma_box_size = sigma_box_size
        self.max_distance = max_distance
        self.use_color_histogram = use_color_histogram
        self.use_bbox = use_bbox
        self.mul = mul
        self.bias = bias
        self.weight_source_sink = weight_source_sink
        self.max_jump = max_jump


def sort_detections(frame):
    """Stable ascending sort by frame index = build_graph's sorted(..., cmp=-a.diff_index(b)) (mot3d/mot3d.py:61-64)."""
    return np.argsort(np.asarray(frame), kind="stable")


def node_weights(conf, spec):
    """weight_confidence_detections: -(confidence*mul + bias) (mot3d/weight_functions.py:37-38)."""
    return -(np.asarray(conf, dtype=np.float64) * spec.mul + spec.bias)


def _bbox_similarity(b1, b2):
    """bbox_size_similarity (mot3d/distance_functions.py:36-45); b = (xmin, ymin, xmax, ymax), broadcast."""
    h1 = b1[..., 3] - b1[..., 1]
    w1 = b1[..., 2] - b1[..., 0]
    h2 = b2[..., 3] - b2[..., 1]
    w2 = b2[..., 2] - b2[..., 0]
    return 1 - (np.abs(h1 - h2) / (np.maximum(h1, h2) + 1e-12) + np.abs(w1 - w2) / (np.maximum(w1, w2) + 1e-12)) / 2


def _hist_similarity(H1, H2):
    """color_histogram_similarity (mot3d/distance_functions.py:21-34); H = [..., C, B], unit channel weights."""
    n_channels = H1.shape[-2]
    scs = 0
    for c in range(n_channels):
        a = H1[..., c, :]
        b = H2[..., c, :]
        sc = np.sum(a * b, axis=-1) / np.sqrt(np.sum(a ** 2, axis=-1) * np.sum(b ** 2, axis=-1))
        scs = scs + 1 * sc
    return scs / n_channels


def _views_terms(views, gi, gj, use_hist, use_bbox):
    """weight_distance_detections_3d's per-view loop (mot3d/weight_functions.py:55-71) for the pairs (gi[k], gj[k]):
    mean over the views of d1 (dict order) that d2 also has of (1 - NCC) and of (1 - bbox similarity).
    views = dict(view_ptr, view_id, bbox [E,4] or None, hist [E,C,B] or None)."""
    vp, vid = views["view_ptr"], views["view_id"]
    mc = np.full(len(gi), np.nan)
    mb = np.full(len(gi), np.nan)
    for k, (i, j) in enumerate(zip(np.asarray(gi).tolist(), np.asarray(gj).tolist())):
        sc, sb = [], []
        for u in range(int(vp[i]), int(vp[i + 1])):
            v = [t for t in range(int(vp[j]), int(vp[j + 1])) if vid[t] == vid[u]]
            if not v:
                continue
            v = v[0]
            if use_hist:
                sc.append(1 - _hist_similarity(views["hist"][u][None], views["hist"][v][None])[0])
            if use_bbox:
                sb.append(1 - _bbox_similarity(views["bbox"][u][None], views["bbox"][v][None])[0])
        if sc:
            mc[k] = np.mean(sc)
        if sb:
            mb[k] = np.mean(sb)
    return mc, mb


def build_transitions(frame, pos, spec, bbox=None, hist=None, views=None):
    """Transition arcs post_i -> pre_j of build_graph's hot loop (mot3d/mot3d.py:106-141) for detections that are
    ALREADY in sorted order: every ordered pair with 0 < frame[j]-frame[i] <= max_jump whose cost is not None
    (weight_distance_detections_2d/_3d, mot3d/weight_functions.py:13-35, 43-75; the 3D variant with a single
    shared view reduces to the same formula). Returns out-CSR: row_ptr[N+1], col[E] ascending per row, w[E] float64.
    Operation order follows the reference: dist = sqrt(((0+dx^2)+dy^2)+dz^2); None if dist > max_distance;
    w = (-exp(-(jump-1)*sigma_jump)) * ((exp(-dist^2/sigma_d^2) * w_hist) * w_bbox)."""
    frame = np.asarray(frame, dtype=np.int64)
    pos = np.asarray(pos, dtype=np.float64)
    n = frame.shape[0]
    lo = np.searchsorted(frame, frame, side="right")  # first j with frame[j] > frame[i]
    hi = np.searchsorted(frame, frame + spec.max_jump, side="right")  # first j with frame[j] > frame[i]+max_jump
    rows, cols, ws = [], [], []
    sd2 = spec.sigma_distance ** 2
    # group sources by frame so that each group shares one candidate range
    starts = np.flatnonzero(np.r_[True, frame[1:] != frame[:-1]])
    ends = np.r_[starts[1:], n]
    for s, e in zip(starts.tolist(), ends.tolist()):
        a, b = int(lo[s]), int(hi[s])
        if b <= a:
            continue
        d = pos[s:e, None, :] - pos[None, a:b, :]
        acc = np.zeros((e - s, b - a))
        for k in range(pos.shape[1]):
            acc = acc + d[:, :, k] ** 2
        dist = np.sqrt(acc)
        keep = ~(dist > spec.max_distance)
        ii, jj = np.nonzero(keep)  # row-major: ascending i then ascending j -> reference emission order
        if ii.size == 0:
            continue
        dsel = dist[ii, jj]
        prod = np.exp(-dsel ** 2 / sd2)
        gi = ii + s
        gj = jj + a
        if views is not None and (spec.use_color_histogram or spec.use_bbox):
            # 3-D detections: both terms averaged over the shared 2-D views (mot3d/weight_functions.py:55-71)
            mc, mb = _views_terms(views, gi, gj, spec.use_color_histogram, spec.use_bbox)
            if spec.use_color_histogram:
                prod = prod * np.exp(-mc ** 2 / spec.sigma_color_histogram ** 2)
            if spec.use_bbox:
                prod = prod * np.exp(-mb ** 2 / spec.sigma_box_size ** 2)
        else:
            if spec.use_color_histogram:
                chs = 1 - _hist_similarity(hist[gi], hist[gj])
                prod = prod * np.exp(-chs ** 2 / spec.sigma_color_histogram ** 2)
            if spec.use_bbox:
                bss = 1 - _bbox_similarity(bbox[gi], bbox[gj])
                prod = prod * np.exp(-bss ** 2 / spec.sigma_box_size ** 2)
        jump = frame[gj] - frame[gi]
        w = -np.exp(-(jump - 1) * spec.sigma_jump) * prod
        rows.append(gi)
        cols.append(gj)
        ws.append(w)
    if rows:
        rows = np.concatenate(rows)
        cols = np.concatenate(cols)
        ws = np.concatenate(ws)
    else:
        rows = np.zeros(0, dtype=np.int64)
        cols = np.zeros(0, dtype=np.int64)
        ws = np.zeros(0, dtype=np.float64)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, rows + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    return row_ptr, cols.astype(np.int64), ws


def graph_arcs(n, entry_w, node_w, exit_w, row_ptr, col, w, entry_mask=None, exit_mask=None):
    """Arc list (1-based tail, head, weight) in the order graph_to_text emits it for a graph made by build_graph
    (mot3d/mot3d.py:75-104,218-226): S=1, pre_m=2+2m, post_m=3+2m, T=2n+2; first every S->pre_m; then per detection
    pre_m->post_m, post_m->T, post_m->pre_j (j ascending). entry_w/exit_w may be scalars or per-detection arrays."""
    n = int(n)
    entry_w = np.broadcast_to(np.asarray(entry_w), (n,))
    exit_w = np.broadcast_to(np.asarray(exit_w), (n,))
    node_w = np.asarray(node_w)
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    if entry_mask is None:
        entry_mask = np.ones(n, dtype=bool)
    if exit_mask is None:
        exit_mask = np.ones(n, dtype=bool)
    T = 2 * n + 2
    m = np.arange(n)
    deg = np.diff(row_ptr)
    per_det = 1 + exit_mask.astype(np.int64) + deg
    n_entry = int(entry_mask.sum())
    E = n_entry + int(per_det.sum())
    tail = np.empty(E, dtype=np.int32)
    head = np.empty(E, dtype=np.int32)
    ww = np.empty(E, dtype=np.asarray(w).dtype if len(w) else node_w.dtype)
    tail[:n_entry] = 1
    head[:n_entry] = 2 + 2 * m[entry_mask]
    ww[:n_entry] = entry_w[entry_mask]
    base = n_entry + np.r_[0, np.cumsum(per_det)[:-1]]
    tail[base] = 2 + 2 * m
    head[base] = 3 + 2 * m
    ww[base] = node_w
    ex_idx = base[exit_mask] + 1
    tail[ex_idx] = 3 + 2 * m[exit_mask]
    head[ex_idx] = T
    ww[ex_idx] = exit_w[exit_mask]
    if len(col):
        src = np.repeat(m, deg)
        within = np.arange(len(col)) - np.repeat(row_ptr[:-1], deg)
        pos = np.repeat(base + 1 + exit_mask.astype(np.int64), deg) + within
        tail[pos] = 3 + 2 * src
        head[pos] = 2 + 2 * np.asarray(col)
        ww[pos] = w
    return tail, head, ww


def arcs_to_text(n_nodes, tail, head, w):
    """graph_to_text (mot3d/mot3d.py:218-226) from arc arrays with float weights."""
    lines = ["p min {} {}".format(n_nodes, len(tail))]
    for s, t, x in zip(tail.tolist(), head.tolist(), np.asarray(w, dtype=np.float64).tolist()):
        lines.append("a {} {} {:0.7f}".format(s, t, x))
    return lines


def arcs_to_text_units(n_nodes, tail, head, c):
    lines = ["p min {} {}".format(n_nodes, len(tail))]
    for s, t, x in zip(tail.tolist(), head.tolist(), np.asarray(c).tolist()):
        lines.append("a {} {} {}".format(s, t, units_to_text(x)))
    return lines


def text_sha(lines):
    """First 16 hex of SHA-256 of "\n".join(lines)+"\n" (SURVEY.md 8c)."""
    txt = "\n".join(lines) + "\n"
    return hashlib.sha256(txt.encode()).hexdigest()[:16]



# ------------------------------------------------------------------------------------------- tracklet linking
class TrackletSpec(object):
    """Parameters of weight_distance_tracklets_2d/_3d (mot3d/weight_functions.py:80-84, 128-132)."""

    def __init__(self, kind="tracklets_2d", sigma_color_histogram=0.3, sigma_motion=50, alpha=0.7, cutoff_motion=0.1,
                 cutoff_appearance=0.1, max_distance=None, use_color_histogram=True, mul=1, bias=0,
                 weight_source_sink=1, max_jump=12):
        self.kind = kind
        self.sigma_color_histogram = sigma_color_histogram
        self.sigma_motion = sigma_motion
        self.alpha = alpha
        self.cutoff_motion = cutoff_motion
        self.cutoff_appearance = cutoff_appearance
        self.max_distance = max_distance
        self.use_color_histogram = use_color_histogram
        self.mul = mul
        self.bias = bias
        self.weight_source_sink = weight_source_sink
        self.max_jump = max_jump


def build_tracklet_transitions(d, spec):
    """Arcs between tracklets exactly as build_graph + weight_distance_tracklets_* make them (mot3d/mot3d.py:106-141,
    mot3d/weight_functions.py:80-160, mot3d/distance_functions.py:47-63), from the flat arrays of
    tests/golden/make_golden.py:tracklet_arrays. Returns out-CSR (row_ptr, col, w) in the reference's arc order
    (tracklets in input order). np.polyfit is called once per end instead of once per pair (same fit)."""
    hp, tp = d["t_head_ptr"], d["t_tail_ptr"]
    n = len(hp) - 1
    hidx, tidx, hpos, tpos = d["t_head_idx"], d["t_tail_idx"], d["t_head_pos"], d["t_tail_pos"]
    fit_h = [np.polyfit(hidx[hp[i]:hp[i + 1]], hpos[hp[i]:hp[i + 1]], 1) for i in range(n)]
    fit_t = [np.polyfit(tidx[tp[i]:tp[i + 1]], tpos[tp[i]:tp[i + 1]], 1) for i in range(n)]

    def deviation(polys, idx2, pts2):
        pred = np.vstack([np.polyval(q, idx2) for q in polys.T]).T
        return np.linalg.norm(pred - pts2, axis=1).mean()

    has_hist = d.get("t_has_hist")
    rows, cols, ws = [], [], []
    for i in range(n):
        head_index = hidx[hp[i + 1] - 1]
        for j in range(n):
            if i == j:
                continue
            jump = tidx[tp[j]] - head_index
            if not (0 < jump <= spec.max_jump):
                continue
            if spec.max_distance is not None:
                dist = np.sqrt(sum((a - b) ** 2 for a, b in zip(hpos[hp[i + 1] - 1], tpos[tp[j]])))
                if dist > spec.max_distance:
                    continue
            if spec.kind == "tracklets_2d" and d["t_tail_access"][j]:
                continue
            dev = deviation(fit_h[i], tidx[tp[j]:tp[j + 1]], tpos[tp[j]:tp[j + 1]]) + \
                deviation(fit_t[j], hidx[hp[i]:hp[i + 1]], hpos[hp[i]:hp[i + 1]])
            wm = np.exp(-dev ** 2 / spec.sigma_motion ** 2)
            if wm < spec.cutoff_motion:
                continue
            if spec.use_color_histogram and spec.kind == "tracklets_3d" and "t_view_ptr" in d:
                # mean over the views both tracklets have (t1's dict order) of 1 - NCC(head of t1, tail of t2)
                # (mot3d/weight_functions.py:146-158)
                vp, vid = d["t_view_ptr"], d["t_view_id"]
                sims = []
                for u in range(int(vp[i]), int(vp[i + 1])):
                    v = [t for t in range(int(vp[j]), int(vp[j + 1])) if vid[t] == vid[u]]
                    if v:
                        sims.append(1 - _hist_similarity(d["t_view_head_hist"][u], d["t_view_tail_hist"][v[0]]))
                wa = np.exp(-np.mean(sims) ** 2 / spec.sigma_color_histogram ** 2)
                if wa < spec.cutoff_appearance:
                    continue
                w = -((1 - spec.alpha) * wm + spec.alpha * wa)
            elif spec.use_color_histogram and has_hist is not None and has_hist[i] and has_hist[j]:
                chs = 1 - _hist_similarity(d["t_head_hist"][i], d["t_tail_hist"][j])
                wa = np.exp(-chs ** 2 / spec.sigma_color_histogram ** 2)
                if wa < spec.cutoff_appearance:
                    continue
                w = -((1 - spec.alpha) * wm + spec.alpha * wa)
            else:
                w = -wm
            rows.append(i)
            cols.append(j)
            ws.append(w)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, np.asarray(rows, dtype=np.int64) + 1, 1)
    return np.cumsum(row_ptr), np.asarray(cols, dtype=np.int64), np.asarray(ws, dtype=np.float64)

# ------------------------------------------------------------------------------------------------- input prep
def merge_close_points(points, eps=0.3, return_indexes=False):
    """mot3d/utils/filt.py:6-24 without sklearn: DBSCAN(min_samples=1) = connected components of "distance <= eps",
    clusters in the order of their first member, each replaced by the mean of its members (index order)."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    if len(points) == 0:
        return points
    P = np.asarray(points, dtype=np.float64).reshape(len(points), -1)
    d2 = np.zeros((len(P), len(P)))
    for k in range(P.shape[1]):
        d2 = d2 + (P[:, None, k] - P[None, :, k]) ** 2
    _, lab = connected_components(csr_matrix(d2 <= eps * eps), directed=False)
    first = {}
    for i, l in enumerate(lab.tolist()):
        first.setdefault(l, len(first))
    lab = np.array([first[l] for l in lab.tolist()])
    groups = [np.flatnonzero(lab == c) for c in range(len(first))]
    if return_indexes:
        return [g.tolist() for g in groups]
    out = []
    for g in groups:
        acc = np.zeros(P.shape[1])
        for q in g.tolist():
            acc = acc + P[q]
        out.append(acc / len(g))
    return np.vstack(out)


# ------------------------------------------------------------------------------------------------- solve
def ssp_solve(n_nodes, tail1, head1, cost_units, path_cap=None, batched=False):
    """C restatement on int64 units. tail1/head1 are 1-based as in the text. -> dict(flow, K, path_cost, total).
    batched=True: several paths per search (the rule of the CUDA solver, see ssp_oracle.c); also returns `searches`."""
    tail = np.ascontiguousarray(np.asarray(tail1, dtype=np.int32) - 1)
    head = np.ascontiguousarray(np.asarray(head1, dtype=np.int32) - 1)
    cost = np.ascontiguousarray(cost_units, dtype=np.int64)
    m = tail.shape[0]
    if path_cap is None:
        path_cap = max(1, n_nodes // 2)
    flow = np.zeros(max(m, 1), dtype=np.int32)
    pc = np.zeros(path_cap, dtype=np.int64)
    K = ctypes.c_int32(0)
    total = ctypes.c_int64(0)
    if batched:
        searches = ctypes.c_int32(0)
        rc = _oracle().oracle_ssp_solve_batched(int(n_nodes), int(m), _p(tail, _i32p), _p(head, _i32p),
                                                _p(cost, _i64p), _p(flow, _i32p), ctypes.byref(K), _p(pc, _i64p),
                                                int(path_cap), ctypes.byref(total), ctypes.byref(searches))
        if rc != 0:
            raise RuntimeError("oracle_ssp_solve_batched failed: %d" % rc)
        return dict(flow=flow[:m], K=K.value, path_cost=pc[:K.value].copy(), total=total.value,
                    searches=searches.value)
    rc = _oracle().oracle_ssp_solve(int(n_nodes), int(m), _p(tail, _i32p), _p(head, _i32p), _p(cost, _i64p),
                                    _p(flow, _i32p), ctypes.byref(K), _p(pc, _i64p), int(path_cap),
                                    ctypes.byref(total))
    if rc != 0:
        raise RuntimeError("oracle_ssp_solve failed: %d" % rc)
    return dict(flow=flow[:m], K=K.value, path_cost=pc[:K.value].copy(), total=total.value)


def mussp_solve(n_nodes, tail1, head1, w, path_cap=None):
    """The real muSSP (oracle/_ref) on float weights, arrays in (no text). -> dict(flow, K, path_cost, total, core_s)."""
    tail = np.ascontiguousarray(tail1, dtype=np.int32)
    head = np.ascontiguousarray(head1, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    m = tail.shape[0]
    if path_cap is None:
        path_cap = max(1, n_nodes // 2)
    flow = np.zeros(max(m, 1), dtype=np.int32)
    pc = np.zeros(path_cap, dtype=np.float64)
    K = ctypes.c_int(0)
    total = ctypes.c_double(0)
    core = ctypes.c_double(0)
    rc = _ref().mussp_ref_solve(int(n_nodes), int(m), _p(tail, _i32p), _p(head, _i32p), _p(w, _f64p),
                                _p(flow, _i32p), ctypes.byref(K), _p(pc, _f64p), int(path_cap), ctypes.byref(total),
                                ctypes.byref(core))
    if rc != 0:
        raise RuntimeError("mussp_ref_solve failed: %d" % rc)
    return dict(flow=flow[:m], K=K.value, path_cost=pc[:K.value].copy(), total=total.value, core_s=core.value)


def mussp_solve_text(lines, n_arcs, path_cap=4096):
    """The real muSSP through the reference's text interface (wrapper.cpp:39-88 parse + solve)."""
    text = ("\n".join(lines) + "\n").encode()
    flow = np.zeros(max(n_arcs, 1), dtype=np.int32)
    pc = np.zeros(path_cap, dtype=np.float64)
    K = ctypes.c_int(0)
    total = ctypes.c_double(0)
    core = ctypes.c_double(0)
    parse = ctypes.c_double(0)
    rc = _ref().mussp_ref_solve_text(text, _p(flow, _i32p), int(n_arcs), ctypes.byref(K), _p(pc, _f64p),
                                     int(path_cap), ctypes.byref(total), ctypes.byref(core), ctypes.byref(parse))
    if rc != 0:
        raise RuntimeError("mussp_ref_solve_text failed: %d" % rc)
    return dict(flow=flow[:n_arcs], K=K.value, path_cost=pc[:min(K.value, path_cap)].copy(), total=total.value,
                core_s=core.value, parse_s=parse.value)


def wrapper_solve(lines, verbose=0):
    """Drop-in for the reference's Boost module: wrapper.solve(graph_as_text, verbose) -> [(tail, head), ...]
    (1-based, arcs carrying flow, ascending arc id; wrapper.cpp:139-169)."""
    arcs = [l for l in lines[1:] if l and l[0] == "a"]
    res = mussp_solve_text(lines, len(arcs))
    out = []
    for f, l in zip(res["flow"].tolist(), arcs):
        if f:
            _, s, t, _ = l.split()
            out.append((int(s), int(t)))
    return out


# ------------------------------------------------------------------------------------------ trajectories
def tracks_from_flow(n, tail1, head1, flow):
    """_run_ssp's tail (mot3d/mot3d.py:165-184): S->T paths of the flow sub-graph, kept as detection indices
    (post-nodes, track[::2][1:]), ordered by first node id. Flow is node-disjoint so each path is a chain."""
    S, T = 1, 2 * n + 2
    nxt = {}
    starts = []
    for s, t, f in zip(np.asarray(tail1).tolist(), np.asarray(head1).tolist(), np.asarray(flow).tolist()):
        if not f:
            continue
        if s == S:
            starts.append(t)
        else:
            nxt[s] = t
    tracks = []
    for pre in sorted(starts):
        node = pre
        track = []
        while node != T:
            if node % 2 == 1:  # post node 3+2m
                track.append((node - 3) // 2)
            node = nxt[node]
        tracks.append(track)
    return tracks


def flow_cost(cost_units, flow):
    return int(np.sum(np.asarray(cost_units, dtype=np.int64)[np.asarray(flow) != 0]))
