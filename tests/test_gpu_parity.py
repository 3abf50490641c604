"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes, mot3d_b200/_lib.py), against
 - the golden vectors frozen from the reference itself (tests/golden, see make_golden.py), and
 - the CPU oracle (oracle/) on the same seeded inputs.
Bar: integer work bit-exact (arc sets and order, quantised costs away from rounding ties, total cost, K, path costs,
tracks); float64 weights within 1e-6 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu

RTOL = 1e-6  # float edge weights, BASELINE.json north_star


def _opt(monkeypatch, name, value):
    """Set a library option (mot3d_set_option) for the duration of one test; _reset_options restores the default."""
    from mot3d_b200 import _lib
    _lib.set_option(name, value)
    _PENDING.append(name)


_PENDING = []


@pytest.fixture(autouse=True)
def _reset_options():
    yield
    if _PENDING:
        from mot3d_b200 import _lib
        while _PENDING:
            _lib.set_option(_PENDING.pop(), None)


def _torch():
    import torch
    return torch


def _spec_from(s, mot3d):
    kind = "detections_2d" if s["kind"] == "2d" else "detections_3d"
    dist = {k: s[k] for k in ("sigma_jump", "sigma_distance", "sigma_color_histogram", "sigma_box_size",
                              "max_distance", "use_color_histogram", "use_bbox")}
    return mot3d.CostSpec(kind=kind, distance=dist, confidence=dict(mul=s["mul"], bias=s["bias"]),
                          weight_source_sink=s["weight_source_sink"], max_jump=s["max_jump"])


def _batch_from(d, mot3d, flags=None):
    n = len(d["frame"])
    b, order = mot3d.DetectionBatch.from_arrays(np.array([0, n]), d["frame"], d["pos"], d["conf"], flags=flags,
                                                bbox=d.get("bbox"), hist=d.get("hist"))
    assert (order == np.arange(n)).all()  # fixtures are stored in node order
    v = gu.views_of(d)
    if v is not None:  # 3-D detections with their 2-D views (planar bbox on the device side)
        b.views = dict(view_ptr=v["view_ptr"].astype(np.int32), view_id=v["view_id"].astype(np.int32),
                       bbox=np.ascontiguousarray(v["bbox"].T), hist=np.ascontiguousarray(v["hist"]))
    return b


def _golden_transitions(d):
    """(row_ptr, col, w, cost) of the post_i -> pre_j arcs in reference order, from the golden arc list."""
    n = len(d["frame"])
    tail, head = d["tail"].astype(np.int64), d["head"].astype(np.int64)
    m = (tail % 2 == 1) & (head != 2 * n + 2) & (tail != 1)
    src = (tail[m] - 3) // 2
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, src + 1, 1)
    return np.cumsum(row_ptr), (head[m] - 2) // 2, d["w_float"][m], d["cost"][m]


@pytest.mark.parametrize("pruned,fused", [(True, True), (True, False), (False, False)],
                         ids=["fused", "xwindow", "allpairs"])
@pytest.mark.parametrize("name", gu.DETECTION_CASES)
def test_edge_build_matches_reference(name, pruned, fused):
    import mot3d_b200 as mot3d
    from mot3d_b200 import _lib
    torch = _torch()
    d = gu.load(name)
    spec = _spec_from(d["spec"], mot3d)
    b = _batch_from(d, mot3d).to("cuda")
    row_ptr, col, w, cost = _golden_transitions(d)
    out = mot3d.build_graphs(b, spec, direction=_lib.DIR_OUT, with_weights=True, pruned=pruned, fused=fused)
    assert out.n_edges == len(col)
    assert (out.row_ptr.cpu().numpy() == row_ptr).all()
    e = out.n_edges
    assert (out.col[:e].cpu().numpy() == col).all()                      # same arcs, same order
    np.testing.assert_allclose(out.weight[:e].cpu().numpy(), w, rtol=RTOL, atol=0)
    gu.assert_costs_match(out.cost[:e].cpu().numpy(), cost, w)
    # the solver's orientation is the transpose of the same arc set
    inn = mot3d.build_graphs(b, spec, direction=_lib.DIR_IN, with_weights=True, pruned=pruned, fused=fused)
    assert inn.n_edges == e
    rp = inn.row_ptr.cpu().numpy().astype(np.int64)
    dst = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    srcs = inn.col[:e].cpu().numpy().astype(np.int64)
    o = np.lexsort((dst, srcs))
    assert (srcs[o] == np.repeat(np.arange(len(row_ptr) - 1), np.diff(row_ptr))).all()
    assert (dst[o] == col).all()
    assert (inn.cost[:e].cpu().numpy()[o] == out.cost[:e].cpu().numpy()).all()
    # per-detection arcs
    n = len(d["frame"])
    tail, head = d["tail"], d["head"]
    node_ref = d["cost"][(tail % 2 == 0) & (head == tail + 1)]
    assert (out.node_cost[:n].cpu().numpy() == node_ref).all()
    assert (out.entry_cost[:n].cpu().numpy() == d["cost"][tail == 1]).all()
    assert (out.exit_cost[:n].cpu().numpy() == 0).all()


def _same_graph(a, b):
    e = a.n_edges
    assert b.n_edges == e
    assert (a.row_ptr.cpu() == b.row_ptr.cpu()).all()
    assert (a.col[:e].cpu() == b.col[:e].cpu()).all()
    assert (a.cost[:e].cpu() == b.cost[:e].cpu()).all()
    if a.weight is not None and b.weight is not None:
        assert (a.weight[:e].cpu() == b.weight[:e].cpu()).all()   # bit-identical float64


@pytest.mark.parametrize("cap,acap", [(None, None), (0, None), (None, 0), (0, 0), (64, 256)],
                         ids=["staged", "cands-global", "arcs-global", "all-global", "tiny"])
@pytest.mark.parametrize("name", ["c3_soldiers_pass1", "c4_synth_appearance", "c5_window_f50_p100", "c4_pets_view1_w1"])
def test_fused_build_equals_two_pass_build(name, cap, acap, monkeypatch):
    """k_edge_build_fused (one launch: window search, look-back scan, costs) writes what count -> scan -> fill writes,
    bit for bit, in both orientations - also when a tile's candidates and/or arcs do not fit shared memory and the
    kernel works on the global arrays (forced here by shrinking the staging areas)."""
    import mot3d_b200 as mot3d
    from mot3d_b200 import _lib
    d = gu.load(name)
    spec = _spec_from(d["spec"], mot3d)
    b = _batch_from(d, mot3d).to("cuda")
    if cap is not None:
        _opt(monkeypatch, "edge_cap", cap)
    if acap is not None:
        _opt(monkeypatch, "edge_acap", acap)
    for direction in (_lib.DIR_IN, _lib.DIR_OUT):
        two = mot3d.build_graphs(b, spec, direction=direction, with_weights=True, fused=False)
        one = mot3d.build_graphs(b, spec, direction=direction, with_weights=True, fused=True)
        _same_graph(two, one)


def test_fused_build_row_ranges_and_capacity():
    """Consecutive mot3d_edge_build calls over row ranges append to one CSR through the device-side arc counter;
    a short arc buffer drops the arcs beyond it, raises the flag and still reports exact row offsets."""
    import ctypes
    import mot3d_b200 as mot3d
    from mot3d_b200 import _lib
    from mot3d_b200.batch import _dets_struct, _ptr, _stream
    torch = _torch()
    L = _lib.lib()
    d = gu.load("c5_window_f50_p100")
    spec = _spec_from(d["spec"], mot3d)
    b = _batch_from(d, mot3d).to("cuda")
    ref = mot3d.build_graphs(b, spec, fused=False)
    n, e = b.n, ref.n_edges
    dev = b.frame.device
    cs = spec.to_c()
    spos = torch.empty(b.dim * n, dtype=torch.float64, device=dev)
    sperm = torch.empty(n, dtype=torch.int32, device=dev)
    ds = _dets_struct(b, ref.tkey, spec, spos, sperm)
    _lib.check(L.mot3d_frame_sort(ctypes.byref(ds), _stream(dev)))
    ws_bytes = L.mot3d_edge_build_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)

    def run(cuts, cap):
        row_ptr = torch.full((n + 1,), -7, dtype=torch.int32, device=dev)
        col = torch.full((max(cap, 1),), -1, dtype=torch.int32, device=dev)
        cost = torch.full((max(cap, 1),), -1, dtype=torch.int64, device=dev)
        over = torch.zeros(1, dtype=torch.int32, device=dev)
        counter = torch.zeros(1, dtype=torch.int64, device=dev)
        for k, (a, z) in enumerate(zip(cuts[:-1], cuts[1:])):
            _lib.check(L.mot3d_edge_build(ctypes.byref(ds), ctypes.byref(cs), _lib.DIR_IN, a, z, _ptr(row_ptr), cap,
                                          _ptr(col), _ptr(cost), None, _ptr(over), _ptr(counter) if k else None,
                                          _ptr(counter), _ptr(ws), ws_bytes, _stream(dev)))
        return row_ptr.cpu(), col.cpu(), cost.cpu(), int(over.item()), int(counter.item())

    for cuts in ([0, n], [0, 1, 257, 257, 3000, n], [0, 256, 512, n]):
        row_ptr, col, cost, over, total = run(cuts, e)
        assert over == 0 and total == e
        assert (row_ptr == ref.row_ptr.cpu()).all()
        assert (col[:e] == ref.col[:e].cpu()).all() and (cost[:e] == ref.cost[:e].cpu()).all()
    short = e - 1000
    row_ptr, col, cost, over, total = run([0, n], short)
    assert over == 1 and total == e
    assert (row_ptr == ref.row_ptr.cpu()).all()
    assert (col[:short] == ref.col[:short].cpu()).all() and (cost[:short] == ref.cost[:short].cpu()).all()


def _solve_checks(name, d, sol, n, perm=None):
    from oracle import oracle
    ref = oracle.ssp_solve(int(d["n_nodes"]), d["tail"], d["head"], d["cost"])
    K = int(sol.n_paths[0].item())
    total = int(sol.total_cost[0].item())
    assert total == ref["total"]                         # exact optimum in 1e-7 units
    assert K == ref["K"] == int(d["K"])
    assert (sol.path_cost[:K].cpu().numpy() == ref["path_cost"]).all()
    if name not in gu.MUSSP_BOGUS_COST:
        assert gu.ref_total_units(d) - total == gu.MUSSP_SUBOPTIMAL.get(name, 0)   # vs the real muSSP
    return ref


@pytest.mark.parametrize("name", gu.DETECTION_CASES)
def test_build_solve_extract_matches_mussp(name):
    import mot3d_b200 as mot3d
    from oracle import oracle
    d = gu.load(name)
    spec = _spec_from(d["spec"], mot3d)
    b = _batch_from(d, mot3d).to("cuda")
    gb, sol = mot3d.track_batch(b, spec, want_path_costs=True)
    n = len(d["frame"])
    ref = _solve_checks(name, d, sol, n)
    tracks = mot3d.tracks_to_lists(np.array([0, n]), sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(),
                                   sol.track_det.cpu().numpy())[0]
    ref_tracks = oracle.tracks_from_flow(n, d["tail"], d["head"], ref["flow"])
    assert tracks == ref_tracks                                       # oracle (exact optimum)
    if gu.MUSSP_SUBOPTIMAL.get(name, 0) == 0:                         # and the reference's own trajectories
        gold = [d["track_det"][d["track_ptr"][k]:d["track_ptr"][k + 1]].tolist()
                for k in range(len(d["track_ptr"]) - 1)]
        assert tracks == gold
    # flow consistency: next/prev are inverse of each other, tracks cover exactly the used detections
    nxt, prv = sol.next[:n].cpu().numpy(), sol.prev[:n].cpu().numpy()
    for i in np.flatnonzero(nxt >= 0).tolist():
        assert prv[nxt[i]] == i
    assert sorted(i for t in tracks for i in t) == np.flatnonzero(prv != -1).tolist()


@pytest.mark.parametrize("name", ["wrapper_test_graph", "c2_tracklets_pass2", "c3_soldiers_pass2", "mussp_seq07"])
def test_solver_on_reference_graphs(name):
    """Solver-only vectors: the wrapper's own smoke graph, the two tracklet-linking graphs (not frame sorted,
    backward-numbered arcs) and muSSP's sample graph (20 210 detections: the global-memory state path)."""
    import mot3d_b200 as mot3d
    from mot3d_b200 import dimacs
    d = gu.load(name)
    gb, perm = dimacs.graph_from_arcs(int(d["n_nodes"]), d["tail"], d["head"], d["cost"])
    sol = mot3d.solve_graphs(gb, want_path_costs=True)
    ref = _solve_checks(name, d, sol, len(perm))
    # same flow as muSSP (these optima are unique): successor of every detection
    n = len(perm)
    nxt = sol.next[:n].cpu().numpy()
    tail, head, flow = d["tail"].astype(np.int64), d["head"].astype(np.int64), ref["flow"]
    T = int(d["n_nodes"])
    exp = np.full(n, -1, dtype=np.int64)
    m = (flow != 0) & (tail % 2 == 1) & (tail != 1)
    exp[(tail[m] - 3) // 2] = np.where(head[m] == T, -2, (head[m] - 2) // 2)
    got = np.full(n, -1, dtype=np.int64)
    got[perm] = np.where(nxt >= 0, perm[np.maximum(nxt, 0)], nxt)
    # the flow we return is feasible and costs exactly the reported total, whatever the tie-breaking
    arc_cost = {(int(t), int(h)): int(c) for t, h, c in zip(tail, head, d["cost"])}
    prv = np.full(n, -1, dtype=np.int64)
    prv[perm] = np.where(sol.prev[:n].cpu().numpy() >= 0, perm[np.maximum(sol.prev[:n].cpu().numpy(), 0)],
                         sol.prev[:n].cpu().numpy())
    total = 0
    for i in np.flatnonzero(got != -1).tolist():
        total += arc_cost[(2 + 2 * i, 3 + 2 * i)]
        total += arc_cost[(3 + 2 * i, T)] if got[i] == -2 else arc_cost[(3 + 2 * i, 2 + 2 * int(got[i]))]
        if prv[i] == -2:
            total += arc_cost[(1, 2 + 2 * i)]
        else:
            assert got[prv[i]] == i
    assert total == ref["total"]
    if name != "mussp_seq07":  # its optimum is not unique (equal-cost alternatives); the others are
        assert (got == exp).all()


def test_public_api_drop_in():
    """The reference's own call sequence (examples/simple_2d/example.py:41-75) on the C1 vector, through
    build_graph / solve_graph with lambdas around the built-in costs; text identical to the reference's."""
    import mot3d_b200 as mot3d
    import mot3d_b200.weight_functions as wf
    from oracle import oracle
    d = gu.load("c1_simple_2d")
    rng = np.random.RandomState(1)
    n = len(d["frame"])
    shuffled = rng.permutation(n)
    hist = d["hist"]
    dets_sorted = [mot3d.Detection2D(int(d["frame"][i]), d["pos"][i], confidence=float(d["conf"][i]),
                                     color_histogram=[hist[i, 0], hist[i, 1], hist[i, 2]], bbox=list(d["bbox"][i]))
                   for i in range(n)]
    # any input order: equal frames keep their relative order (stable sort), so shuffle whole frames only
    frames = np.unique(d["frame"])
    dets = [dets_sorted[i] for f in rng.permutation(frames) for i in np.flatnonzero(d["frame"] == f)]
    weight_distance = lambda d1, d2: wf.weight_distance_detections_2d(d1, d2, sigma_jump=1, sigma_distance=10,
                                                                      sigma_color_histogram=0.3, sigma_box_size=0.3,
                                                                      max_distance=15, use_color_histogram=True,
                                                                      use_bbox=True)
    weight_confidence = lambda d: wf.weight_confidence_detections_2d(d, mul=1, bias=0)
    g = mot3d.build_graph(dets, weight_source_sink=0.1, max_jump=4, verbose=False,
                          weight_confidence=weight_confidence, weight_distance=weight_distance)
    assert g is not None and len(g) == int(d["n_nodes"])
    assert [x.pre_node for x in dets_sorted] == [2 + 2 * m for m in range(n)]
    assert [x.post_node for x in dets_sorted] == [3 + 2 * m for m in range(n)]
    lines = mot3d.graph_to_text(g)
    assert oracle.text_sha(lines) == str(d["sha"])
    trajectories = mot3d.solve_graph(g, verbose=False, method="muSSP")
    got = [[dets_sorted.index(x) for x in t] for t in trajectories]
    gold = [d["track_det"][d["track_ptr"][k]:d["track_ptr"][k + 1]].tolist() for k in range(len(d["track_ptr"]) - 1)]
    assert got == gold
    # networkx view used by plot_graph (reference mot3d/utils/vis.py:78-113)
    assert g.nodes[2]["label"] == "pre-node" and g.nodes[3]["detection"] is dets_sorted[0]
    assert len(g.edges()) == len(d["tail"]) and g.has_edge(1, 2)
    with pytest.raises(ValueError):
        mot3d.solve_graph(g, verbose=False, method="nope")
    with pytest.raises(NotImplementedError):
        mot3d.solve_graph(g, verbose=False, method="SSP")
    assert mot3d.build_graph([], verbose=False) is None
    with pytest.raises(TypeError):
        mot3d.build_graph(dets, verbose=False, weight_distance=lambda a, b: -1.0)
    # a single pair evaluated on the GPU
    w = wf.weight_distance_detections_2d(dets_sorted[0], dets_sorted[2], sigma_jump=1, sigma_distance=10,
                                         max_distance=15, use_color_histogram=False, use_bbox=False)
    tail, head = d["tail"], d["head"]
    k = np.flatnonzero((tail == 3) & (head == 6))
    assert (w is None and len(k) == 0) or abs(w - d["w_float"][k[0]]) <= RTOL * abs(d["w_float"][k[0]])


def test_public_api_multiview_3d_detections():
    """Detection3D objects that carry their 2-D views ({view: Detection2D}, reference mot3d/types.py:40-45) through
    build_graph / solve_graph with weight_distance_detections_3d(use_color_histogram=True, use_bbox=True): both
    appearance terms averaged over the shared views in the tail detection's dict order (reference
    mot3d/weight_functions.py:55-71). Text and trajectories are the reference's; a pair without a common view is an
    error here (the reference computes nan)."""
    import mot3d_b200 as mot3d
    import mot3d_b200.weight_functions as wf
    d = gu.load("c3_synth_multiview")
    n = len(d["frame"])
    vp, vid, vb, vh = d["view_ptr"], d["view_id"], d["view_bbox"], d["view_hist"]

    def views(i, drop=None):
        out = {}
        for u in range(int(vp[i]), int(vp[i + 1])):
            if drop is not None and int(vid[u]) in drop:
                continue
            name = "cam%d" % int(vid[u])
            out[name] = mot3d.Detection2D(int(d["frame"][i]), (0.0, 0.0), view=name,
                                          color_histogram=[h for h in vh[u]], bbox=tuple(vb[u]))
        return out

    dets = [mot3d.Detection3D(int(d["frame"][i]), d["pos"][i], confidence=float(d["conf"][i]), detections_2d=views(i))
            for i in range(n)]
    s = d["spec"]
    wd = lambda a, b: wf.weight_distance_detections_3d(a, b, sigma_jump=s["sigma_jump"],
                                                       sigma_distance=s["sigma_distance"],
                                                       sigma_color_histogram=s["sigma_color_histogram"],
                                                       sigma_box_size=s["sigma_box_size"],
                                                       max_distance=s["max_distance"], use_color_histogram=True,
                                                       use_bbox=True)
    wc = lambda x: wf.weight_confidence_detections_3d(x, mul=s["mul"], bias=s["bias"])
    g = mot3d.build_graph(dets, weight_source_sink=s["weight_source_sink"], max_jump=s["max_jump"], verbose=False,
                          weight_confidence=wc, weight_distance=wd)
    tail, head, w = g.arcs()
    assert (tail == d["tail"]).all() and (head == d["head"]).all()
    np.testing.assert_allclose(w, d["w_float"], rtol=RTOL, atol=0)
    trajectories = mot3d.solve_graph(g, verbose=False)
    got = [[dets.index(x) for x in t] for t in trajectories]
    gold = [d["track_det"][d["track_ptr"][k]:d["track_ptr"][k + 1]].tolist() for k in range(len(d["track_ptr"]) - 1)]
    assert got == gold
    # one pair through the same kernels
    k = int(np.flatnonzero((d["tail"] % 2 == 1) & (d["tail"] != 1) & (d["head"] != 2 * n + 2))[0])
    i, j = (int(d["tail"][k]) - 3) // 2, (int(d["head"][k]) - 2) // 2
    w1 = wf.weight_distance_detections_3d(dets[i], dets[j], sigma_jump=s["sigma_jump"],
                                          sigma_distance=s["sigma_distance"],
                                          sigma_color_histogram=s["sigma_color_histogram"],
                                          sigma_box_size=s["sigma_box_size"], max_distance=s["max_distance"])
    assert abs(w1 - d["w_float"][k]) <= RTOL * abs(d["w_float"][k])
    # no common view: nan in the reference, an error here
    lonely = [mot3d.Detection3D(0, (0.0, 0.0, 0.0), detections_2d=views(i, drop={0, 1})),
              mot3d.Detection3D(1, (0.1, 0.0, 0.0), detections_2d=views(j, drop={2, 3}))]
    with pytest.raises(ValueError):
        mot3d.build_graph(lonely, max_jump=2, verbose=False, weight_confidence=wc, weight_distance=wd)


def test_batched_graphs_equal_single_graphs():
    """Several different graphs (ragged sizes, one of them empty) in one batch = the same graphs one by one."""
    import mot3d_b200 as mot3d
    names = ["c5_small_f30_p12", "quirk_single_detection", "c5_small_f30_p12", "c5_window_f50_p100"]
    ds = [gu.load(n) for n in names]
    spec = _spec_from(ds[0]["spec"], mot3d)
    sizes = [len(d["frame"]) for d in ds]
    sizes.insert(2, 0)  # an empty graph in the middle
    gp = np.cumsum([0] + sizes)
    frame = np.concatenate([d["frame"] for d in ds])
    pos = np.concatenate([d["pos"] for d in ds])
    conf = np.concatenate([d["conf"] for d in ds])
    b, order = mot3d.DetectionBatch.from_arrays(gp, frame, pos, conf)
    gb, sol = mot3d.track_batch(b.to("cuda"), spec)
    tracks = mot3d.tracks_to_lists(gp, sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(),
                                   sol.track_det.cpu().numpy())
    totals = sol.total_cost.cpu().numpy()
    gi = 0
    for k, sz in enumerate(sizes):
        if sz == 0:
            assert tracks[k] == [] and totals[k] == 0
            continue
        d = ds[gi]
        gi += 1
        if names[gi - 1] == "quirk_single_detection":
            assert tracks[k] == [[0]]  # first path kept although its cost is +0.5 (wrapper.cpp:205,220)
            continue
        assert totals[k] == gu.ref_total_units(d)
        gold = [d["track_det"][d["track_ptr"][t]:d["track_ptr"][t + 1]].tolist()
                for t in range(len(d["track_ptr"]) - 1)]
        assert tracks[k] == gold


def test_host_buffer_entry_point():
    """mot3d_track_batch_host (pinned host buffers in/out, copies inside the call) = the device path."""
    import mot3d_b200 as mot3d
    d = gu.load("c5_small_f30_p12")
    spec = _spec_from(d["spec"], mot3d)
    n = len(d["frame"])
    gp = np.array([0, n, 2 * n])
    b, _ = mot3d.DetectionBatch.from_arrays(gp, np.r_[d["frame"], d["frame"]], np.r_[d["pos"], d["pos"]],
                                            np.r_[d["conf"], d["conf"]])
    ht = mot3d.HostTracker(n_max=2 * n + 7, g_max=3, edge_capacity=20000, dim=2)
    ht.load(b)
    n_edges = ht.run(spec.to_c())
    assert n_edges == 2 * int(((d["tail"] % 2 == 1) & (d["tail"] != 1) & (d["head"] != int(d["n_nodes"]))).sum())
    gold = [d["track_det"][d["track_ptr"][t]:d["track_ptr"][t + 1]].tolist() for t in range(len(d["track_ptr"]) - 1)]
    assert ht.tracks() == [gold, gold]
    assert ht.h_total[:2].tolist() == [gu.ref_total_units(d)] * 2
    # the per-thread streams / events / pinned scratch can be released and come back on the next call
    ht.close()
    assert ht.run(spec.to_c()) == n_edges and ht.tracks() == [gold, gold]
    # too small an arc buffer is reported, not silently truncated
    small = mot3d.HostTracker(n_max=2 * n, g_max=2, edge_capacity=100, dim=2)
    small.load(b)
    with pytest.raises(mot3d.Mot3dError) as e:
        small.run(spec.to_c())
    assert e.value.code == -3


@pytest.mark.parametrize("chunks", [1, 3, 8])
def test_host_entry_point_chunked_pipeline(chunks, monkeypatch):
    """The software pipeline inside mot3d_track_batch_host (chunks of whole graphs on streams of their own, copies
    overlapped with kernels) returns what the device path returns on a ragged batch, whatever the number of chunks;
    a chunk whose share of the arc buffer is too small sends the batch through again in one piece."""
    import mot3d_b200 as mot3d
    rng = np.random.RandomState(11)
    spec = mot3d.CostSpec(kind="detections_2d", max_jump=4, weight_source_sink=1, confidence=dict(mul=1, bias=0),
                          distance=dict(sigma_jump=1, sigma_distance=10, max_distance=30, use_color_histogram=False,
                                        use_bbox=False))
    frames, poss, gp = [], [], [0]
    for g in range(11):
        # ragged: empty graphs, a lone detection, sparse and crowded windows (the crowded ones own most of the arcs)
        F = [0, 1, 12, 25, 3, 40, 0, 18, 30, 9, 22][g]
        P = [0, 1, 6, 14, 2, 30, 0, 4, 3, 25, 8][g]
        extent = 60.0 if g in (5, 9) else 400.0
        pos = rng.rand(P, 2) * extent
        for f in range(F):
            pos = pos + rng.randn(P, 2)
            keep = rng.rand(P) < 0.9
            frames.append(np.full(int(keep.sum()), f))
            poss.append(pos[keep].copy())
        gp.append(sum(len(x) for x in frames))
    frame = np.concatenate(frames)
    pos = np.concatenate(poss)
    gp = np.array(gp)
    n = len(frame)
    conf = 0.5 + 0.5 * rng.rand(n)
    b, order = mot3d.DetectionBatch.from_arrays(gp, frame, pos, conf)
    assert (order == np.arange(n)).all()
    gb, sol = mot3d.track_batch(b.to("cuda"), spec)
    want_tracks = mot3d.tracks_to_lists(gp, sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(),
                                        sol.track_det.cpu().numpy())
    _opt(monkeypatch, "host_chunks", chunks)
    for cap in (gb.n_edges + 64, gb.n_edges):  # roomy, then exact: the crowded chunk overflows its share
        ht = mot3d.HostTracker(n_max=n + 5, g_max=len(gp) + 2, edge_capacity=cap, dim=2)
        ht.load(b)
        assert ht.run(spec.to_c()) == gb.n_edges
        assert ht.tracks() == want_tracks
        assert ht.h_total[:len(gp) - 1].tolist() == sol.total_cost.cpu().tolist()
        assert ht.run(spec.to_c()) == gb.n_edges  # the workspace and the streams are reusable
        assert ht.tracks() == want_tracks
    small = mot3d.HostTracker(n_max=n, g_max=len(gp), edge_capacity=gb.n_edges - 1, dim=2)
    small.load(b)
    with pytest.raises(mot3d.Mot3dError) as e:
        small.run(spec.to_c())
    assert e.value.code == -3


@pytest.mark.parametrize("chunks,pieces", [(1, 1), (1, 4), (2, 3), (3, 8)])
def test_host_entry_point_pieces_with_bbox(chunks, pieces, monkeypatch):
    """The host entry point with the bbox term on (PETS window, use_bbox=True) and the inputs arriving in pieces: the
    per-piece keys / node costs / frame sort / arc build (the build continuing the CSR through the device-side arc
    counter) give what the device path gives, for every chunk x piece split."""
    import mot3d_b200 as mot3d
    d = gu.load("c4_pets_view1_w1")
    spec = _spec_from(d["spec"], mot3d)
    n = len(d["frame"])
    reps = 7
    gp = np.arange(reps + 1) * n
    b, order = mot3d.DetectionBatch.from_arrays(gp, np.tile(d["frame"], reps), np.tile(d["pos"], (reps, 1)),
                                                np.tile(d["conf"], reps), bbox=np.tile(d["bbox"], (reps, 1)))
    assert (order == np.arange(reps * n)).all()
    gb, sol = mot3d.track_batch(b.to("cuda"), spec)
    want = mot3d.tracks_to_lists(gp, sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(),
                                 sol.track_det.cpu().numpy())
    gold = [d["track_det"][d["track_ptr"][t]:d["track_ptr"][t + 1]].tolist() for t in range(len(d["track_ptr"]) - 1)]
    assert want == [gold] * reps
    _opt(monkeypatch, "host_chunks", chunks)
    _opt(monkeypatch, "host_pieces", pieces)
    ht = mot3d.HostTracker(n_max=reps * n, g_max=reps, edge_capacity=gb.n_edges + 8 * reps, dim=2, with_bbox=True)
    ht.load(b)
    assert ht.run(spec.to_c()) == gb.n_edges
    assert ht.tracks() == want
    assert ht.h_total[:reps].tolist() == sol.total_cost.cpu().tolist()


def test_entry_exit_flags_against_oracle():
    """entry_points / exit_point = False (mot3d/mot3d.py:90-99): muSSP cannot run such graphs (UB), the
    integer oracle can."""
    import mot3d_b200 as mot3d
    from mot3d_b200 import _lib
    from oracle import oracle
    d = gu.load("c5_small_f30_p12")
    n = len(d["frame"])
    rng = np.random.RandomState(5)
    flags = (rng.rand(n) < 0.6).astype(np.uint8) * _lib.FLAG_ENTRY | (rng.rand(n) < 0.6).astype(np.uint8) * _lib.FLAG_EXIT
    spec = _spec_from(d["spec"], mot3d)
    b = _batch_from(d, mot3d, flags=flags).to("cuda")
    gb, sol = mot3d.track_batch(b, spec, want_path_costs=True)
    s = d["spec"]
    ospec = oracle.CostSpec(**{k: v for k, v in s.items() if k != "kind"})
    row_ptr, col, w = oracle.build_transitions(d["frame"], d["pos"], ospec)
    tail, head, ww = oracle.graph_arcs(n, 1.0, oracle.node_weights(d["conf"], ospec), 0.0, row_ptr, col, w,
                                       entry_mask=(flags & 1) != 0, exit_mask=(flags & 2) != 0)
    ref = oracle.ssp_solve(2 * n + 2, tail, head, oracle.quantise(ww))
    assert int(sol.total_cost[0].item()) == ref["total"] and int(sol.n_paths[0].item()) == ref["K"]
    tracks = mot3d.tracks_to_lists(np.array([0, n]), sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(),
                                   sol.track_det.cpu().numpy())[0]
    assert tracks == oracle.tracks_from_flow(n, tail, head, ref["flow"])
    for t in tracks:
        assert flags[t[0]] & 1 and flags[t[-1]] & 2


def test_full_size_c5_whole_graph():
    """BASELINE config 5 as ONE graph: 100 000 detections / 1 000 frames / max_jump 8. Size-independent checks:
    arc count, K and total cost equal the real muSSP's (golden summary), rows sorted, every track time-ordered and
    within max_jump / max_distance."""
    import mot3d_b200 as mot3d
    torch = _torch()
    g = gu.load("c5_whole_summary")
    rng = np.random.RandomState(0)
    P, F = 100, 1000
    pos = rng.rand(P, 2) * 1000.0
    vel = rng.randn(P, 2) * 1.0
    frames, positions = [], []
    for f in range(F):
        pos = pos + vel + rng.randn(P, 2) * 0.2
        frames.append(np.full(P, f))
        positions.append(pos.copy())
    frame, xy = np.concatenate(frames), np.concatenate(positions)
    spec = mot3d.CostSpec(kind="detections_2d",
                          distance=dict(sigma_jump=1, sigma_distance=10, max_distance=30, use_color_histogram=False,
                                        use_bbox=False),
                          confidence=dict(mul=1, bias=0), weight_source_sink=1, max_jump=8)
    b, _ = mot3d.DetectionBatch.from_arrays(np.array([0, len(frame)]), frame, xy, np.full(len(frame), 0.9))
    gb, sol = mot3d.track_batch(b.to("cuda"), spec)
    assert gb.n_edges == int(g["n_transitions"])
    assert int(sol.n_paths[0].item()) == int(g["K"])
    assert int(sol.total_cost[0].item()) == int(g["total_units"])
    rp = gb.row_ptr.cpu().numpy()
    col = gb.col[:gb.n_edges].cpu().numpy()
    seg = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    same = seg[1:] == seg[:-1]
    assert (col[1:][same] > col[:-1][same]).all()          # sources ascend inside every row
    tracks = mot3d.tracks_to_lists(np.array([0, len(frame)]), sol.track_count.cpu().numpy(),
                                   sol.track_ptr.cpu().numpy(), sol.track_det.cpu().numpy())[0]
    assert [t[0] for t in tracks] == sorted(t[0] for t in tracks)
    for t in tracks:
        t = np.array(t)
        dt = np.diff(frame[t])
        assert (dt > 0).all() and (dt <= 8).all()
        assert (np.linalg.norm(np.diff(xy[t], axis=0), axis=1) <= 30).all()


# ----------------------------------------------------------------------------------------------- tracklet linking
def _tracklet_spec_from(s, mot3d):
    dist = {k: s[k] for k in ("sigma_color_histogram", "sigma_motion", "alpha", "cutoff_motion", "cutoff_appearance",
                              "max_distance", "use_color_histogram")}
    return mot3d.TrackletCostSpec(kind=s["kind"], distance=dist, confidence=dict(mul=s["mul"], bias=s["bias"]),
                                  weight_source_sink=s["weight_source_sink"], max_jump=s["max_jump"])


def _packed_from_fixture(d, dim):
    n = len(d["t_conf"])
    out = dict(n=n, dim=dim, head_ptr=d["t_head_ptr"].astype(np.int32), tail_ptr=d["t_tail_ptr"].astype(np.int32),
               head_idx=d["t_head_idx"].astype(np.int32), tail_idx=d["t_tail_idx"].astype(np.int32),
               head_pos=np.ascontiguousarray(d["t_head_pos"].reshape(-1, dim).T),
               tail_pos=np.ascontiguousarray(d["t_tail_pos"].reshape(-1, dim).T), conf=d["t_conf"],
               tail_access=d["t_tail_access"].astype(np.uint8), flags=None)
    if "t_view_ptr" in d:  # 3-D tracklets with per-view histograms
        out.update(head_hist=None, tail_hist=None, view_ptr=d["t_view_ptr"].astype(np.int32),
                   view_id=d["t_view_id"].astype(np.int32), view_head_hist=d["t_view_head_hist"],
                   view_tail_hist=d["t_view_tail_hist"])
    elif "t_head_hist" in d:
        out.update(head_hist=d["t_head_hist"], tail_hist=d["t_tail_hist"],
                   head_has_hist=d["t_has_hist"].astype(np.uint8), tail_has_hist=d["t_has_hist"].astype(np.uint8))
    return out


@pytest.mark.parametrize("name", gu.TRACKLET_CASES)
def test_tracklet_build_and_solve_match_reference(name):
    """Row a18: weight_distance_tracklets_2d/_3d + linear_motion on the GPU, then the same solver."""
    import mot3d_b200 as mot3d
    from mot3d_b200 import tracklets as trk, spec as mspec
    d = gu.load(name)
    tspec = _tracklet_spec_from(d["spec"], mot3d)
    dim = tspec.dim
    packed = _packed_from_fixture(d, dim)
    row_ptr, col, w, cost = trk.build_tracklet_arcs(packed, tspec)
    # golden transitions
    n = packed["n"]
    tail, head = d["tail"].astype(np.int64), d["head"].astype(np.int64)
    m = (tail % 2 == 1) & (tail != 1) & (head != 2 * n + 2)
    src = (tail[m] - 3) // 2
    rp = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rp, src + 1, 1)
    assert (row_ptr == np.cumsum(rp)).all()
    assert (col == (head[m] - 2) // 2).all()                     # same arcs, same order
    np.testing.assert_allclose(w, d["w_float"][m], rtol=RTOL, atol=0)
    gu.assert_costs_match(cost, d["cost"][m], d["w_float"][m])
    # solve on the reference's own quantised arc costs (bit-identical input), compare with muSSP
    s = d["spec"]
    en = np.full(n, mspec.quantise_scalar(s["weight_source_sink"]), dtype=np.int64)
    node_cost = d["cost"][(tail % 2 == 0) & (head == tail + 1)]
    tracks, total, K, _ = trk.solve_tracklet_graph(n, row_ptr, col, d["cost"][m], en, node_cost,
                                                   np.zeros(n, dtype=np.int64))
    assert total == gu.ref_total_units(d) and K == int(d["K"])
    gold = [d["track_det"][d["track_ptr"][k]:d["track_ptr"][k + 1]].tolist() for k in range(len(d["track_ptr"]) - 1)]
    assert tracks == gold


@pytest.mark.parametrize("name", ["c2_tracklets_pass2", "c3_soldiers_pass2"])
def test_tracklet_public_api_drop_in(name):
    """build_graph / solve_graph on DetectionTracklet2D/3D objects, as examples/simple_2d_tracklets/example.py:78-108
    and examples/soldiers_3d/example.py:81-101 call them."""
    import mot3d_b200 as mot3d
    import mot3d_b200.weight_functions as wf
    from oracle import oracle
    d = gu.load(name)
    s = d["spec"]
    n = len(d["t_conf"])
    hp = d["t_head_ptr"]
    tracklets = []
    for i in range(n):
        idx = d["t_head_idx"][hp[i]:hp[i + 1]]       # window=None: the head window is the whole tracklet
        pts = d["t_head_pos"][hp[i]:hp[i + 1]]
        if s["kind"] == "tracklets_2d":
            hist = [h for h in d["t_head_hist"][i]] if "t_head_hist" in d else None
            chain = [mot3d.Detection2D(int(f), p, color_histogram=hist, bbox=[0, 0, 1, 1]) for f, p in zip(idx, pts)]
            tracklets.append(mot3d.DetectionTracklet2D(chain))
        else:
            chain = [mot3d.Detection3D(int(f), p) for f, p in zip(idx, pts)]
            tracklets.append(mot3d.DetectionTracklet3D(chain))
    kw = {k: s[k] for k in ("sigma_color_histogram", "sigma_motion", "alpha", "cutoff_motion", "cutoff_appearance",
                            "max_distance", "use_color_histogram")}
    if s["kind"] == "tracklets_2d":
        wd = lambda t1, t2: wf.weight_distance_tracklets_2d(t1, t2, **kw)
        wc = lambda t: wf.weight_confidence_tracklets_2d(t, mul=s["mul"], bias=s["bias"])
    else:
        wd = lambda t1, t2: wf.weight_distance_tracklets_3d(t1, t2, **kw)
        wc = lambda t: wf.weight_confidence_tracklets_3d(t, mul=s["mul"], bias=s["bias"])
    g = mot3d.build_graph(tracklets, weight_source_sink=s["weight_source_sink"], max_jump=s["max_jump"], verbose=False,
                          weight_confidence=wc, weight_distance=wd)
    assert g is not None and len(g) == int(d["n_nodes"])
    tail, head, w = g.arcs()
    assert (tail == d["tail"]).all() and (head == d["head"]).all()
    np.testing.assert_allclose(w, d["w_float"], rtol=RTOL, atol=0)
    if name == "c2_tracklets_pass2":
        assert oracle.text_sha(mot3d.graph_to_text(g)) == str(d["sha"])
    trajectories = mot3d.solve_graph(g, verbose=False)
    got = [[tracklets.index(t) for t in tr] for tr in trajectories]
    gold = [d["track_det"][d["track_ptr"][k]:d["track_ptr"][k + 1]].tolist() for k in range(len(d["track_ptr"]) - 1)]
    assert got == gold
    assert g.total_cost == gu.ref_total_units(d)


def test_tracklet_public_api_multiview_3d():
    """build_graph / solve_graph on DetectionTracklet3D objects whose per-view 2-D tracklets carry histograms:
    weight_distance_tracklets_3d(use_color_histogram=True) averages the appearance term over the shared views
    (reference mot3d/weight_functions.py:146-158). Arcs, weights, total and trajectories are the reference's."""
    import types as pytypes
    import mot3d_b200 as mot3d
    import mot3d_b200.weight_functions as wf
    d = gu.load("c3_synth_multiview_tracklets")
    s = d["spec"]
    n = len(d["t_conf"])
    hp, tp, vp = d["t_head_ptr"], d["t_tail_ptr"], d["t_view_ptr"]
    tracklets = []
    for i in range(n):
        t = object.__new__(mot3d.DetectionTracklet3D)   # the ends as the reference's constructor leaves them
        t._init_item(float(d["t_conf"][i]), None)
        t.head = mot3d.Detection3D(int(d["t_head_idx"][hp[i + 1] - 1]), d["t_head_pos"][hp[i + 1] - 1])
        t.head.indexes, t.head.positions = d["t_head_idx"][hp[i]:hp[i + 1]], d["t_head_pos"][hp[i]:hp[i + 1]]
        t.tail = mot3d.Detection3D(int(d["t_tail_idx"][tp[i]]), d["t_tail_pos"][tp[i]])
        t.tail.indexes, t.tail.positions = d["t_tail_idx"][tp[i]:tp[i + 1]], d["t_tail_pos"][tp[i]:tp[i + 1]]
        t.head.access_point = t.tail.access_point = False
        t.tracklets_2d = {}
        for u in range(int(vp[i]), int(vp[i + 1])):
            t.tracklets_2d["cam%d" % int(d["t_view_id"][u])] = pytypes.SimpleNamespace(
                head=pytypes.SimpleNamespace(color_histogram=d["t_view_head_hist"][u]),
                tail=pytypes.SimpleNamespace(color_histogram=d["t_view_tail_hist"][u]))
        t.tracklet = [t.tail, t.head]
        tracklets.append(t)
    kw = {k: s[k] for k in ("sigma_color_histogram", "sigma_motion", "alpha", "cutoff_motion", "cutoff_appearance",
                            "max_distance", "use_color_histogram")}
    wd = lambda t1, t2: wf.weight_distance_tracklets_3d(t1, t2, **kw)
    wc = lambda t: wf.weight_confidence_tracklets_3d(t, mul=s["mul"], bias=s["bias"])
    g = mot3d.build_graph(tracklets, weight_source_sink=s["weight_source_sink"], max_jump=s["max_jump"], verbose=False,
                          weight_confidence=wc, weight_distance=wd)
    assert g is not None and len(g) == int(d["n_nodes"])
    tail, head, w = g.arcs()
    assert (tail == d["tail"]).all() and (head == d["head"]).all()
    np.testing.assert_allclose(w, d["w_float"], rtol=RTOL, atol=0)
    trajectories = mot3d.solve_graph(g, verbose=False)
    got = [[tracklets.index(t) for t in tr] for tr in trajectories]
    gold = [d["track_det"][d["track_ptr"][k]:d["track_ptr"][k + 1]].tolist() for k in range(len(d["track_ptr"]) - 1)]
    assert got == gold
    assert g.total_cost == gu.ref_total_units(d)
    # no common view between two linkable tracklets: nan in the reference, an error here
    for t in tracklets:
        t.tracklets_2d = {("a" if tracklets.index(t) % 2 else "b"): next(iter(t.tracklets_2d.values()))}
    with pytest.raises(ValueError):
        mot3d.build_graph(tracklets, weight_source_sink=s["weight_source_sink"], max_jump=s["max_jump"],
                          verbose=False, weight_confidence=wc, weight_distance=wd)


# ------------------------------------------------------------------------------------------------------ fuzzing
def test_random_graphs_against_oracle():
    """300 random small windows in ONE batch (ragged sizes incl. empty and single-detection graphs, missing frames,
    dense and sparse scenes, entry costs that make the first path unprofitable) vs the integer oracle: K, total cost
    and a feasible flow of exactly that cost for every graph. Exercises the component split, the first-path rule and
    the work queue on shapes the golden vectors do not cover."""
    import mot3d_b200 as mot3d
    from mot3d_b200 import _lib
    from oracle import oracle
    rng = np.random.RandomState(123)
    G = 300
    frames, poss, confs, sizes = [], [], [], []
    for g in range(G):
        n = int(rng.choice([0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144]))
        F = int(rng.randint(1, 30))
        extent = float(rng.choice([20.0, 60.0, 200.0]))
        f = np.sort(rng.randint(0, F, size=n))
        frames.append(f)
        poss.append(rng.rand(n, 2) * extent)
        confs.append(rng.rand(n))
        sizes.append(n)
    gp = np.cumsum([0] + sizes)
    kw = dict(sigma_jump=0.5, sigma_distance=8, max_distance=25, use_color_histogram=False, use_bbox=False)
    for wss, mul, bias in [(0.3, 1, 0), (2.5, 1, 0.1), (0.05, 0.5, 0)]:
        spec = mot3d.CostSpec(kind="detections_2d", distance=kw, confidence=dict(mul=mul, bias=bias),
                              weight_source_sink=wss, max_jump=3)
        b, order = mot3d.DetectionBatch.from_arrays(gp, np.concatenate(frames), np.concatenate(poss),
                                                    np.concatenate(confs))
        gb, sol = mot3d.track_batch(b.to("cuda"), spec, want_path_costs=True)
        K = sol.n_paths.cpu().numpy()
        total = sol.total_cost.cpu().numpy()
        path_cost = sol.path_cost.cpu().numpy()
        tracks = mot3d.tracks_to_lists(gp, sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(),
                                       sol.track_det.cpu().numpy())
        ospec = oracle.CostSpec(weight_source_sink=wss, max_jump=3, mul=mul, bias=bias, **kw)
        for g in range(G):
            n = sizes[g]
            if n == 0:
                assert K[g] == 0 and total[g] == 0 and tracks[g] == []
                continue
            fr, xy, cf = b.frame[gp[g]:gp[g + 1]], b.pos[:, gp[g]:gp[g + 1]].T, b.conf[gp[g]:gp[g + 1]]
            row_ptr, col, w = oracle.build_transitions(fr, xy, ospec)
            tail, head, ww = oracle.graph_arcs(n, float(wss), oracle.node_weights(cf, ospec), 0.0, row_ptr, col, w)
            cost = oracle.quantise(ww)
            ref = oracle.ssp_solve(2 * n + 2, tail, head, cost)
            assert K[g] == ref["K"], (g, n, K[g], ref["K"])
            assert total[g] == ref["total"], (g, n, total[g], ref["total"])
            # path costs in the order the reference finds them, the rest of the graph's slice is "no path"
            assert (path_cost[gp[g]:gp[g] + K[g]] == ref["path_cost"]).all(), g
            assert (path_cost[gp[g] + K[g]:gp[g + 1]] == _lib.INF_COST).all(), g
            # our tracks form a feasible flow of exactly that cost
            arc = {(int(t), int(h)): int(c) for t, h, c in zip(tail, head, cost)}
            got = 0
            used = set()
            for t in tracks[g]:
                got += arc[(1, 2 + 2 * t[0])] + arc[(3 + 2 * t[-1], 2 * n + 2)]
                for a_, b_ in zip(t[:-1], t[1:]):
                    got += arc[(3 + 2 * a_, 2 + 2 * b_)]
                for d_ in t:
                    assert d_ not in used
                    used.add(d_)
                    got += arc[(2 + 2 * d_, 3 + 2 * d_)]
            assert got == total[g] and len(tracks[g]) == K[g]
            assert [t[0] for t in tracks[g]] == sorted(t[0] for t in tracks[g])


def test_two_pass_pipeline_drop_in():
    """examples/simple_2d_tracklets/example.py end to end through the drop-in surface: detections -> build_graph ->
    solve_graph -> split_trajectory_modulo / remove_short_trajectories -> DetectionTracklet2D -> build_graph ->
    solve_graph. Both passes must reproduce the reference's graphs (sha of graph_to_text) and trajectories."""
    import mot3d_b200 as mot3d
    import mot3d_b200.weight_functions as wf
    from oracle import oracle
    d1, d2 = gu.load("c1_simple_2d"), gu.load("c2_tracklets_pass2")
    n = len(d1["frame"])
    hist = d1["hist"]
    dets = [mot3d.Detection2D(int(d1["frame"][i]), d1["pos"][i], confidence=float(d1["conf"][i]),
                              color_histogram=[hist[i, 0], hist[i, 1], hist[i, 2]], bbox=list(d1["bbox"][i]))
            for i in range(n)]
    g = mot3d.build_graph(dets, weight_source_sink=0.1, max_jump=4, verbose=False,
                          weight_confidence=lambda d: wf.weight_confidence_detections_2d(d, mul=1, bias=0),
                          weight_distance=lambda a, b: wf.weight_distance_detections_2d(
                              a, b, sigma_jump=1, sigma_distance=10, sigma_color_histogram=0.3, sigma_box_size=0.3,
                              max_distance=15, use_color_histogram=True, use_bbox=True))
    trajectories = mot3d.solve_graph(g, verbose=False)
    tracklets = []
    for traj in trajectories:
        tracklets += mot3d.split_trajectory_modulo(traj, length=5)
    tracklets = mot3d.remove_short_trajectories(tracklets, th_length=2)
    dts = [mot3d.DetectionTracklet2D(t) for t in tracklets]
    assert len(dts) == len(d2["t_conf"])
    g2 = mot3d.build_graph(dts, weight_source_sink=0.1, max_jump=20, verbose=False,
                           weight_confidence=lambda t: wf.weight_confidence_tracklets_2d(t, mul=1, bias=0),
                           weight_distance=lambda a, b: wf.weight_distance_tracklets_2d(
                               a, b, max_distance=None, sigma_color_histogram=0.3, sigma_motion=50, alpha=0.7,
                               cutoff_motion=0.1, cutoff_appearance=0.1, use_color_histogram=True))
    assert oracle.text_sha(mot3d.graph_to_text(g2)) == str(d2["sha"])
    final = mot3d.solve_graph(g2, verbose=False)
    got = [[dts.index(t) for t in tr] for tr in final]
    gold = [d2["track_det"][d2["track_ptr"][k]:d2["track_ptr"][k + 1]].tolist()
            for k in range(len(d2["track_ptr"]) - 1)]
    assert got == gold
    joined = [mot3d.concat_tracklets([t.tracklet for t in tr]) for tr in final]
    assert [len(j) for j in joined] == [13, 13]
    for j in joined:
        assert [d.index for d in j] == sorted(d.index for d in j)


@pytest.mark.parametrize("case", ["c4_online_loop", "c4_online_loop_hist"])
def test_online_window_loop_matches_reference(case):
    """BASELINE config 4 (PETS-shaped online sliding-window tracking): mot3d_b200.online.Tracker2D / Scene, the mirror
    of the reference's loop (projects/PETS2009S2L1/singleview_tracking.py:76-134, tracking_routines.py:16-52,
    scene.py:8-52), over 4 batches of 50 frames of C4-synth. After every batch the scene - ids, frames, interpolated
    positions and boxes of the active and the completed trajectories - must be what the reference's own loop left
    (tests/golden/c4_online_loop.npz, made by running that loop unmodified)."""
    import json
    import mot3d_b200 as mot3d
    from mot3d_b200.online import Scene, Tracker2D
    d = gu.load(case)
    cfg = json.loads(str(d["config"]))
    scene = Scene(cfg["completed_after"])
    tracker = Tracker2D(scene, None, cfg, image_size=(768, 576))
    ptr, det = d["det_ptr"], d["det"]
    hist = d["det_hist"] if "det_hist" in d else None   # (_hist: both costs use their appearance terms)
    B = cfg["batch_size"]
    n_frames = len(ptr) - 1
    for b in range(n_frames // B):
        dets = [[mot3d.Detection2D(f, det[r, :2].copy(), confidence=float(det[r, 2]), bbox=tuple(det[r, 3:7]),
                                   color_histogram=None if hist is None else hist[r].copy())
                 for r in range(int(ptr[f]), int(ptr[f + 1]))] for f in range(b * B, (b + 1) * B)]
        tracker.update((b + 1) * B, dets)
        for name, store in (("active", scene.active), ("completed", scene.completed)):
            ids = sorted(store)
            assert ids == d["b%d_%s_ids" % (b, name)].tolist(), (b, name)
            assert np.cumsum([0] + [len(store[i]) for i in ids]).tolist() == d["b%d_%s_ptr" % (b, name)].tolist()
            assert [x.index for i in ids for x in store[i]] == d["b%d_%s_index" % (b, name)].tolist()
            pos = np.array([np.asarray(x.position, dtype=np.float64) for i in ids for x in store[i]]).reshape(-1, 2)
            box = np.array([np.asarray(x.bbox, dtype=np.float64) for i in ids for x in store[i]]).reshape(-1, 4)
            np.testing.assert_allclose(pos, d["b%d_%s_pos" % (b, name)], rtol=1e-12, atol=1e-9)
            np.testing.assert_allclose(box, d["b%d_%s_bbox" % (b, name)], rtol=1e-12, atol=1e-9)
    assert len(scene.tracklets) == int(d["n_tracklets"])
    assert len(scene.completed) == (5 if hist is not None else 2) and len(scene.active) == 6


def _scene_arrays(store):
    ids = sorted(store)
    ptr = np.cumsum([0] + [len(store[i][0]) for i in ids]).tolist()
    cat = lambda k, w: (np.concatenate([store[i][k] for i in ids]) if ids else np.zeros((0, w))).reshape(-1, w)
    return ids, ptr, cat(0, 1).reshape(-1).astype(np.int64).tolist(), cat(1, 2), cat(2, 4)


@pytest.mark.parametrize("case", ["c4_online_loop", "c4_online_loop_hist"])
def test_online_loop_with_device_resident_scene_matches_reference(case):
    """DeviceTracker2D keeps the active trajectories in device memory between updates and does the
    cutting, windowing, linking, concatenation and hole filling in kernels (csrc/twopass.cu: k_split_tracks with carried
    trajectories, k_online_concat). Same golden as the object loop: after every batch of C4-synth the ids, frames,
    interpolated positions and boxes of the active and the completed trajectories are what the reference's own loop
    left (tests/golden/c4_online_loop.npz). _hist: the detections carry colour histograms and both costs use them - the
    scene keeps a histogram per point and the tracklet cost compares window medians (k_window_median_hist)."""
    import json
    from mot3d_b200.online_device import DeviceTracker2D
    d = gu.load(case)
    hist = d["det_hist"] if "det_hist" in d else None
    hs = lambda lo, hi: None if hist is None else hist[lo:hi]
    cfg = json.loads(str(d["config"]))
    tracker = DeviceTracker2D(cfg, completed_after=cfg["completed_after"])
    ptr, det = d["det_ptr"], d["det"]
    B = cfg["batch_size"]
    n_frames = len(ptr) - 1
    for b in range(n_frames // B):
        lo, hi = int(ptr[b * B]), int(ptr[(b + 1) * B])
        frame = np.repeat(np.arange(b * B, (b + 1) * B), np.diff(ptr[b * B:(b + 1) * B + 1]))
        tracker.update((b + 1) * B, frame, det[lo:hi, :2], det[lo:hi, 2], det[lo:hi, 3:7], hs(lo, hi))
        for name, store in (("active", tracker.active()), ("completed", tracker.completed)):
            ids, p_, index, pos, box = _scene_arrays(store)
            assert ids == d["b%d_%s_ids" % (b, name)].tolist(), (b, name)
            assert p_ == d["b%d_%s_ptr" % (b, name)].tolist()
            assert index == d["b%d_%s_index" % (b, name)].tolist()
            np.testing.assert_allclose(pos, d["b%d_%s_pos" % (b, name)].reshape(-1, 2), rtol=1e-12, atol=1e-9)
            np.testing.assert_allclose(box, d["b%d_%s_bbox" % (b, name)].reshape(-1, 4), rtol=1e-12, atol=1e-9)
    assert tracker.n_tracklets_seen == int(d["n_tracklets"])
    assert len(tracker.completed) == (5 if hist is not None else 2) and len(tracker.ids) == 6
    # the detection pass runs in preallocated buffers sized by a guess: too small an arc buffer is noticed, the buffers
    # are re-sized from the build's own count and the batch runs again - same scene
    from mot3d_b200.pipeline import TrackPipeline
    again = DeviceTracker2D(cfg, completed_after=cfg["completed_after"])
    again._pipe = TrackPipeline(1024, 1, 64, dim=2)
    for b in range(n_frames // B):
        lo, hi = int(ptr[b * B]), int(ptr[(b + 1) * B])
        frame = np.repeat(np.arange(b * B, (b + 1) * B), np.diff(ptr[b * B:(b + 1) * B + 1]))
        again.update((b + 1) * B, frame, det[lo:hi, :2], det[lo:hi, 2], det[lo:hi, 3:7], hs(lo, hi))
    assert again._pipe.edge_capacity > 64
    a, b_ = _scene_arrays(tracker.active()), _scene_arrays(again.active())
    assert a[:3] == b_[:3] and (a[3] == b_[3]).all() and (a[4] == b_[4]).all()
    assert again.n_tracklets_seen == tracker.n_tracklets_seen and sorted(again.completed) == sorted(tracker.completed)


def test_online_loop_device_scene_equals_object_loop_on_a_long_sequence():
    """The same two loops (objects: mot3d_b200.online.Tracker2D, pinned to the reference above; arrays on the device:
    DeviceTracker2D) over 12 batches of a sequence where people enter, leave for good and come back after gaps, with
    some empty batches: trajectories complete, ids are never reused, carried trajectories are extended, left alone or
    linked behind another one. Every detection carries a colour histogram and the tracklet cost uses it: the windows of
    carried trajectories mix points with a histogram and filled points without. Scenes must agree after every batch."""
    import mot3d_b200 as mot3d
    from mot3d_b200.online import Scene, Tracker2D
    from mot3d_b200.online_device import DeviceTracker2D
    cfg = dict(batch_size=20, completed_after=30,
               detections=dict(weight_source_sink=1, max_jump=4, verbose=False,
                               weight_distance=dict(sigma_distance=10, sigma_jump=4, sigma_color_histogram=0.1,
                                                    sigma_box_size=0.1, max_distance=30, use_color_histogram=False,
                                                    use_bbox=True),
                               weight_confidence=dict(mul=0, bias=0)),
               tracklets=dict(weight_source_sink=0.1, max_jump=40, length_endpoints=5, verbose=False,
                              weight_distance=dict(sigma_color_histogram=0.3, sigma_motion=100, alpha=0.2,
                                                   cutoff_motion=0.1, cutoff_appearance=0.1, use_color_histogram=True),
                              weight_confidence=dict(mul=0, bias=0.11)))
    rng = np.random.RandomState(5)
    B, n_batches = cfg["batch_size"], 12
    people = []
    for k in range(14):
        t0 = int(rng.randint(0, 160))
        people.append(dict(t0=t0, t1=t0 + int(rng.randint(25, 110)), p=rng.uniform(50, 700, 2), v=rng.uniform(-3, 3, 2),
                           gap=(int(rng.randint(t0 + 10, t0 + 60)), int(rng.randint(4, 25))), wh=rng.uniform(20, 60, 2)))
    rows, hists = [], []
    for q in people:
        q["h"] = rng.normal(0, 1, (3, 6))
    for f in range(B * n_batches):
        if 100 <= f < 120:
            continue  # an empty batch
        for q in people:
            if not (q["t0"] <= f < q["t1"]) or (q["gap"][0] <= f < q["gap"][0] + q["gap"][1]) or rng.rand() < 0.08:
                continue
            c = q["p"] + q["v"] * (f - q["t0"]) + rng.normal(0, 0.7, 2)
            w, h = q["wh"] * (1 + rng.normal(0, 0.02, 2))
            rows.append((f, c[0], c[1], rng.uniform(0.5, 1), c[0] - w / 2, c[1] - h / 2, c[0] + w / 2, c[1] + h / 2))
            # a histogram per detection: the person's pattern + noise; half the people change clothes after a gap
            flip = -1.0 if (people.index(q) % 2 and f >= q["gap"][0]) else 1.0
            hists.append(flip * q["h"] + rng.normal(0, 0.3, (3, 6)))
    rows, hists = np.array(rows), np.array(hists)
    class Region(object):  # valid_region: everything but a strip on the left
        @staticmethod
        def is_inside(p):
            return p[0] > 90

    scene = Scene(cfg["completed_after"])
    t_obj = Tracker2D(scene, Region, cfg, image_size=(768, 576))
    t_dev = DeviceTracker2D(cfg, completed_after=cfg["completed_after"], valid_region=Region)
    n_done = 0
    for b in range(n_batches):
        m = (rows[:, 0] >= b * B) & (rows[:, 0] < (b + 1) * B)
        sel, hsel = rows[m], hists[m]
        dets = [[mot3d.Detection2D(int(r[0]), r[1:3].copy(), confidence=float(r[3]), bbox=tuple(r[4:8]),
                                   color_histogram=hh.copy())
                 for r, hh in zip(sel, hsel) if int(r[0]) == f] for f in range(b * B, (b + 1) * B)]
        assert t_obj.update((b + 1) * B, dets) == 0   # (no tracklet is skipped at the second valid_region test)
        t_dev.update((b + 1) * B, sel[:, 0].astype(np.int64), sel[:, 1:3], sel[:, 3], sel[:, 4:8], hsel)
        for name, want, got in (("active", scene.active, t_dev.active()), ("completed", scene.completed, t_dev.completed)):
            assert sorted(want) == sorted(got), (b, name)
            for i in want:
                assert [x.index for x in want[i]] == got[i][0].tolist(), (b, name, i)
                np.testing.assert_allclose(np.array([np.asarray(x.position, dtype=np.float64) for x in want[i]]),
                                           got[i][1], rtol=1e-12, atol=1e-9)
                np.testing.assert_allclose(np.array([np.asarray(x.bbox, dtype=np.float64) for x in want[i]]),
                                           got[i][2], rtol=1e-12, atol=1e-9)
        n_done = len(scene.completed)
    assert n_done >= 3 and scene.id_last == t_dev.id_last and len(scene.tracklets) == t_dev.n_tracklets_seen


def test_single_track_pass_equals_full_solver(monkeypatch):
    """k_single_track finishes most components of crowded windows straight from the first tree; what it writes must
    be what the staged solver writes for the same components. 60 C5 windows (5 000 detections each, ~88 components
    per window), solved with the pass on and off: same flow, same K, same totals, same path costs."""
    import mot3d_b200 as mot3d
    from bench import make_workload, DIST_KW, MAX_JUMP
    gp, frame, pos, conf = make_workload(3, seed0=77)
    spec = mot3d.CostSpec(kind="detections_2d", distance=DIST_KW, confidence=dict(mul=1, bias=0),
                          weight_source_sink=1, max_jump=MAX_JUMP)
    b, _ = mot3d.DetectionBatch.from_arrays(gp, frame, pos, conf)
    db = b.to("cuda")
    gb = mot3d.build_graphs(db, spec)
    sol_fast = mot3d.solve_graphs(gb, want_path_costs=True)
    stats = sol_fast.stats.cpu().numpy().reshape(-1, mot3d._lib.SOLVE_STATS)
    assert stats[:, 24].sum() > 0.5 * 60 * 80          # the pass did take most components
    _opt(monkeypatch, "solve_no_fast", 1)
    sol_full = mot3d.solve_graphs(gb, want_path_costs=True)
    assert sol_full.stats.cpu().numpy().reshape(-1, mot3d._lib.SOLVE_STATS)[:, 24].sum() == 0
    n = db.n
    assert (sol_fast.next[:n].cpu() == sol_full.next[:n].cpu()).all()
    assert (sol_fast.prev[:n].cpu() == sol_full.prev[:n].cpu()).all()
    assert (sol_fast.n_paths.cpu() == sol_full.n_paths.cpu()).all()
    assert (sol_fast.total_cost.cpu() == sol_full.total_cost.cpu()).all()
    assert (sol_fast.path_cost[:n].cpu() == sol_full.path_cost[:n].cpu()).all()
    assert int(sol_fast.n_paths.sum().item()) == 60 * 100


def test_merge_close_points_matches_reference():
    """mot3d_merge_close_points (csrc/filt.cu) vs the reference's merge_close_points (mot3d/utils/filt.py:6-24) on the
    350 frames of examples/soldiers_3d (all frames in one launch) and on a crowded 400-point set with chains of
    neighbours (cluster membership and order, means bit for bit)."""
    import mot3d_b200 as mot3d
    from mot3d_b200.utils.filt import merge_close_points, merge_close_points_batch
    d = gu.load("c3_merge_close_points")
    rp, mp = d["raw_ptr"], d["merged_ptr"]
    sets = [d["raw"][rp[f]:rp[f + 1]] for f in range(len(rp) - 1)]
    got = merge_close_points_batch(sets, eps=0.25)
    for f in range(len(sets)):
        assert (np.asarray(got[f]).reshape(-1, 3) == d["merged"][mp[f]:mp[f + 1]]).all()
    # the positions the 3-D golden vector was built from are exactly these
    c3 = gu.load("c3_soldiers_pass1")
    assert (np.concatenate([np.asarray(g).reshape(-1, 3) for g in got]) == c3["pos"]).all()
    groups = merge_close_points(d["crowd"], eps=0.3, return_indexes=True)
    gp = d["crowd_group_ptr"]
    assert [len(g) for g in groups] == np.diff(gp).tolist()
    assert [i for g in groups for i in g] == d["crowd_group_idx"].tolist()
    assert (merge_close_points(d["crowd"], eps=0.3) == d["crowd_merged"]).all()
    assert merge_close_points([], eps=0.3) == []


def test_components_against_scipy():
    """mot3d_components (csrc/components.cu): the regrouping of every graph by connected component of its transition
    arcs, checked against scipy on a ragged batch; repeated, because the union-find runs lock-free."""
    import ctypes
    import torch
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components
    import mot3d_b200 as mot3d
    from mot3d_b200 import _lib
    rng = np.random.RandomState(5)
    sizes = [0, 1, 37, 3000, 5000, 800, 8192, 2, 9000, 4999] + [int(s) for s in rng.randint(1, 4000, size=40)]
    frames, poss = [], []
    for s in sizes:
        f = np.sort(rng.randint(0, 40, size=s))
        frames.append(f)
        poss.append(rng.rand(s, 2) * (300.0 + 5.0 * np.sqrt(s)))
    gp = np.cumsum([0] + sizes).astype(np.int32)
    frame, pos = np.concatenate(frames), np.concatenate(poss)
    spec = mot3d.CostSpec(kind="detections_2d", max_jump=3, weight_source_sink=1,
                          distance=dict(sigma_jump=1, sigma_distance=10, max_distance=25, use_color_histogram=False,
                                        use_bbox=False))
    b, _ = mot3d.DetectionBatch.from_arrays(gp, frame, pos, np.full(len(frame), 0.9))
    gb = mot3d.build_graphs(b.to("cuda"), spec)
    n, G, E = len(frame), len(sizes), gb.n_edges
    L = _lib.lib()
    dev = gb.row_ptr.device
    i32 = dict(dtype=torch.int32, device=dev)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    row_ptr, col = gb.row_ptr.cpu().numpy(), gb.col.cpu().numpy()[:E]
    rows = np.repeat(np.arange(n), np.diff(row_ptr))
    ws_bytes = L.mot3d_components_workspace_bytes(n, G)
    for rep in range(8):
        perm, inv = torch.full((n,), -1, **i32), torch.full((n,), -1, **i32)
        cptr, cgraph = torch.full((n + 1,), -1, **i32), torch.full((n,), -1, **i32)
        cbase, ncg = torch.full((G,), -1, **i32), torch.full((4,), -1, **i32)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(L.mot3d_components(G, n, ptr(gb.dets.graph_ptr), ptr(gb.row_ptr), ptr(gb.col), E, ptr(perm),
                                      ptr(inv), ptr(cptr), ptr(cgraph), ptr(cbase), ptr(ncg), ptr(ws), ws_bytes,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        perm, inv, cptr, cgraph, cbase = [t.cpu().numpy() for t in (perm, inv, cptr, cgraph, cbase)]
        n_cg = int(ncg[0].item())
        assert (perm[inv] == np.arange(n)).all() and cptr[n_cg] == n
        c = 0
        for g, s in enumerate(sizes):
            g0 = gp[g]
            if s == 0:
                continue
            assert cbase[g] == c
            if s > 8192:  # kept whole
                assert cgraph[c] == g and cptr[c] == g0 and (perm[g0:g0 + s] == np.arange(g0, g0 + s)).all()
                c += 1
                continue
            m = (rows >= g0) & (rows < g0 + s)
            A = sp.coo_matrix((np.ones(m.sum()), (rows[m] - g0, col[m] - g0)), shape=(s, s))
            k, lab = connected_components(A, directed=False)
            first = np.full(k, s)
            np.minimum.at(first, lab, np.arange(s))
            rank = np.argsort(np.argsort(first))[lab]        # components ordered by their first detection
            order = np.lexsort((np.arange(s), rank))         # grouped by component, ascending detection inside
            assert (perm[g0:g0 + s] == g0 + order).all()
            starts = g0 + np.concatenate([[0], np.cumsum(np.bincount(rank, minlength=k))[:-1]])
            assert (cptr[c:c + k] == starts).all() and (cgraph[c:c + k] == g).all()
            c += k
        assert c == n_cg


@pytest.mark.gpu
def test_track_payload_pack_unpack_on_device():
    """mot3d_pack_tracks == its numpy mirror bit for bit; mot3d_unpack_tracks restores count / ptr / det."""
    torch = _torch()
    import mot3d_b200 as mot3d
    from mot3d_b200 import _lib
    d = gu.load("c5_window_f50_p100")
    spec = _spec_from(d["spec"], mot3d)
    b1 = _batch_from(d, mot3d)
    # two graphs in one batch (the second one shifted, so that its tracks differ)
    gp = np.array([0, b1.n, 2 * b1.n], dtype=np.int32)
    hb, _ = mot3d.DetectionBatch.from_arrays(gp, np.concatenate([b1.frame, b1.frame]),
                                             np.concatenate([b1.pos.T, b1.pos.T + 0.25]), np.concatenate([b1.conf, b1.conf]))
    db = hb.to("cuda")
    gb, sol = mot3d.track_batch(db, spec)
    n, G = db.n, db.n_graphs
    L = _lib.lib()
    payload = torch.full((int(L.mot3d_track_payload_words(n, G)),), -1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.mot3d_pack_tracks(G, db.graph_ptr.data_ptr(), sol.track_count.data_ptr(), sol.track_ptr.data_ptr(),
                                   sol.track_det.data_ptr(), payload.data_ptr(), st))
    want = mot3d.pack_tracks_host(gp, sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(),
                                  sol.track_det.cpu().numpy())
    assert np.array_equal(payload.cpu().numpy(), want)
    cnt = torch.zeros(G, dtype=torch.int32, device="cuda")
    ptr = torch.zeros(n + G + 1, dtype=torch.int32, device="cuda")
    det = torch.zeros(n, dtype=torch.int32, device="cuda")
    _lib.check(L.mot3d_unpack_tracks(G, db.graph_ptr.data_ptr(), payload.data_ptr(), cnt.data_ptr(), ptr.data_ptr(),
                                     det.data_ptr(), st))
    from mot3d_b200.batch import tracks_to_lists
    got = tracks_to_lists(gp, cnt.cpu().numpy(), ptr.cpu().numpy(), det.cpu().numpy())
    assert got == tracks_to_lists(gp, sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(), sol.track_det.cpu().numpy())
    assert got == mot3d.unpack_tracks_host(gp, want)
    assert sum(len(t) for t in got[0]) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["p2p", "collective"])
def test_two_gpu_sharded_solve_and_gather(transport):
    """Two ranks, one GPU each: shard_graphs over a ragged mix, TrackPipeline per rank, TrackGather to rank 0, compared
    with one GPU solving everything (tests/gather_worker.py). Skipped on a single-GPU box."""
    torch = _torch()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import os
    import subprocess
    import sys
    port = 33500 + os.getpid() % 2000 + (7 if transport == "p2p" else 0)
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gather_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(port), "gpu", transport], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK transport=%s" % transport in outs[0], outs


def test_key_preconditions_are_reported():
    """Time keys must be non-decreasing inside a graph and fit 31 bits over the batch (csrc/edge_build.cu: k_make_keys);
    both are reported by the device path and by the host-buffer call instead of producing wrong arcs."""
    import mot3d_b200 as mot3d
    d = gu.load("c5_small_f30_p12")
    spec = _spec_from(d["spec"], mot3d)
    n = len(d["frame"])
    dev = _torch().device("cuda")
    # (1) frames of a graph out of order, sorting switched off
    fr = d["frame"].copy()
    fr[[3, n - 4]] = fr[[n - 4, 3]]
    b, _ = mot3d.DetectionBatch.from_arrays(np.array([0, n]), fr, d["pos"], d["conf"], sort=False)
    with pytest.raises(ValueError):
        mot3d.build_graphs(b.to(dev), spec)
    ht = mot3d.HostTracker(n_max=n, g_max=1, edge_capacity=20000, dim=2)
    ht.load(b)
    with pytest.raises(mot3d.Mot3dError) as e:
        ht.run(spec.to_c())
    assert e.value.code == -1
    # (2) sparse frame indices: three graphs spanning 2^30 frames each do not fit the batch's 31-bit key space
    fr = d["frame"].astype(np.int64).copy()
    fr[fr == fr.max()] = 1 << 30
    b3, _ = mot3d.DetectionBatch.from_arrays(np.array([0, n, 2 * n, 3 * n]), np.r_[fr, fr, fr],
                                             np.r_[d["pos"], d["pos"], d["pos"]], np.r_[d["conf"], d["conf"], d["conf"]])
    with pytest.raises(mot3d.Mot3dError) as e:
        mot3d.build_graphs(b3.to(dev), spec)
    assert e.value.code == -4
    ht3 = mot3d.HostTracker(n_max=3 * n, g_max=3, edge_capacity=60000, dim=2)
    ht3.load(b3)
    with pytest.raises(mot3d.Mot3dError) as e:
        ht3.run(spec.to_c())
    assert e.value.code == -4
    # one such graph alone is fine, and equals the graph with the gap closed up to max_jump + 1
    b1, _ = mot3d.DetectionBatch.from_arrays(np.array([0, n]), fr, d["pos"], d["conf"])
    g1 = mot3d.build_graphs(b1.to(dev), spec)
    assert g1.n_edges > 0


def test_graph_file_to_solution(tmp_path):
    """N4 end to end at file level: graph_to_text -> save_graph -> file -> dimacs.parse_text -> graph_from_arcs -> CUDA
    solve: K and total of the reference for a detection graph built here, its golden, and the adversarial a9 graph."""
    import mot3d_b200 as mot3d
    from mot3d_b200 import dimacs
    for name in ("c1_simple_2d", "a9_clipped_arcs", "c5_small_f30_p12"):
        d = gu.load(name)
        spec = _spec_from(d["spec"], mot3d)
        b = _batch_from(d, mot3d).to("cuda")
        out = mot3d.build_graphs(b, spec, direction=mot3d._lib.DIR_OUT, with_weights=True)
        n = len(d["frame"])
        from oracle import oracle
        row_ptr = out.row_ptr.cpu().numpy().astype(np.int64)
        tail, head, ww = oracle.graph_arcs(n, float(d["spec"]["weight_source_sink"]),
                                           -(d["conf"] * d["spec"]["mul"] + d["spec"]["bias"]), 0.0, row_ptr,
                                           out.col[:out.n_edges].cpu().numpy().astype(np.int64),
                                           out.weight[:out.n_edges].cpu().numpy())
        path = str(tmp_path / (name + ".txt"))
        mot3d.save_graph(oracle.arcs_to_text(2 * n + 2, tail, head, ww), path)
        n_nodes, t2, h2, c2 = dimacs.parse_text(open(path).read().splitlines())
        assert n_nodes == int(d["n_nodes"]) and (t2 == d["tail"]).all() and (h2 == d["head"]).all()
        gu.assert_costs_match(c2, d["cost"], d["w_float"])
        gb, perm = dimacs.graph_from_arcs(n_nodes, t2, h2, d["cost"])
        sol = mot3d.solve_graphs(gb)
        assert int(sol.n_paths[0].item()) == int(d["K"])
        assert int(sol.total_cost[0].item()) == gu.ref_total_units(d)


@pytest.mark.parametrize("name", ["c2_tracklets_pass2", "c3_soldiers_pass2"])
def test_tracklet_costs_on_real_pairs(name):
    """weight_distance_tracklets_2d/_3d(t1, t2, ...) called on two real tracklets evaluates that pair on the GPU (the
    two-tracklet arc build): the reference's arc weights of the tracklet fixtures, None where the reference has no
    arc."""
    import mot3d_b200 as mot3d
    import mot3d_b200.weight_functions as wf
    d = gu.load(name)
    s = d["spec"]
    n = len(d["t_conf"])
    hp = d["t_head_ptr"]
    tracklets = []
    for i in range(n):
        idx, pts = d["t_head_idx"][hp[i]:hp[i + 1]], d["t_head_pos"][hp[i]:hp[i + 1]]  # window=None: the whole tracklet
        if s["kind"] == "tracklets_2d":
            hist = [h for h in d["t_head_hist"][i]] if "t_head_hist" in d else None
            tracklets.append(mot3d.DetectionTracklet2D([mot3d.Detection2D(int(f), p, color_histogram=hist,
                                                                          bbox=[0, 0, 1, 1]) for f, p in zip(idx, pts)]))
        else:
            tracklets.append(mot3d.DetectionTracklet3D([mot3d.Detection3D(int(f), p) for f, p in zip(idx, pts)]))
    kw = {k: s[k] for k in ("sigma_color_histogram", "sigma_motion", "alpha", "cutoff_motion", "cutoff_appearance",
                            "max_distance", "use_color_histogram")}
    fn = wf.weight_distance_tracklets_2d if s["kind"] == "tracklets_2d" else wf.weight_distance_tracklets_3d
    tail, head = d["tail"].astype(np.int64), d["head"].astype(np.int64)
    m = np.flatnonzero((tail % 2 == 1) & (tail != 1) & (head != int(d["n_nodes"])))
    arcs = {((int(tail[e]) - 3) // 2, (int(head[e]) - 2) // 2): float(d["w_float"][e]) for e in m}
    assert len(arcs) >= 2
    for (i, j), w in list(arcs.items())[:8]:
        got = fn(tracklets[i], tracklets[j], **kw)
        assert got is not None and abs(got - w) <= RTOL * abs(w), (i, j, got, w)
    # ordered pairs in time without an arc in the reference's graph: None (cut-offs / max_distance), within max_jump
    miss = 0
    for i in range(n):
        for j in range(n):
            jump = tracklets[j].tail.index - tracklets[i].head.index
            if 0 < jump <= s["max_jump"] and (i, j) not in arcs and miss < 4:
                assert fn(tracklets[i], tracklets[j], **kw) is None
                miss += 1
    with pytest.raises(AssertionError):  # the reference's linear_motion asserts the order in time
        fn(tracklets[1], tracklets[1], **kw)


@pytest.mark.parametrize("name", ["c4_synth_appearance", "c3_synth_multiview", "c5_small_f30_p12"])
def test_track_pipeline_with_appearance_costs(name):
    """TrackPipeline (the scheduler's per-rank path: preallocated buffers, no host synchronisation) with the appearance
    terms: colour histograms + boxes of 2-D detections, per-view terms of 3-D detections. Same arcs, costs, totals and
    tracks as build_graphs / solve_graphs, i.e. the reference's."""
    import mot3d_b200 as mot3d
    from mot3d_b200.pipeline import TrackPipeline
    d = gu.load(name)
    spec = _spec_from(d["spec"], mot3d)
    n = len(d["frame"])
    b = _batch_from(d, mot3d).to("cuda")
    gb, sol = mot3d.track_batch(b, spec)
    pipe = TrackPipeline(n, 1, gb.n_edges + 16, dim=b.dim)
    for _ in range(2):  # (buffers are reused from run to run)
        pipe.run(b, spec)
    pipe.check_overflow()
    E = gb.n_edges
    assert pipe.n_edges() == E
    assert (pipe.col[:E] == gb.col[:E]).all() and (pipe.cost[:E] == gb.cost[:E]).all()
    assert int(pipe.total_cost[0].item()) == int(sol.total_cost[0].item()) == gu.ref_total_units(d)
    gold = [d["track_det"][d["track_ptr"][t]:d["track_ptr"][t + 1]].tolist() for t in range(len(d["track_ptr"]) - 1)]
    assert pipe.tracks(np.array([0, n]))[0] == gold
    # too small an arc buffer: reported, and the appearance pass stays inside it
    small = TrackPipeline(n, 1, max(E // 2, 1), dim=b.dim)
    small.run(b, spec)
    with pytest.raises(mot3d.Mot3dError):
        small.check_overflow()


@pytest.mark.parametrize("case", ["c3_soldiers_3d", "c5_small_2d_window", "c5_small_2d_window_hist"])
def test_device_resident_two_pass_equals_host_glue(case):
    """link_tracks (pass 1 -> tracklets -> pass 2 -> trajectories on the device, csrc/twopass.cu) against the drop-in
    path on Python objects (split_trajectory_modulo / remove_short_trajectories / DetectionTracklet / build_graph /
    solve_graph / concat_tracklets), which is pinned to the reference (test_two_pass_pipeline_drop_in and the pass-2
    fixtures). C3 is examples/soldiers_3d/example.py: 3-D detections, length 10, th_length 2, whole tracklets at both ends;
    the 2-D case uses windows of 3 frames."""
    import mot3d_b200 as mot3d
    import mot3d_b200.weight_functions as wf
    if case == "c3_soldiers_3d":
        d = gu.load("c3_soldiers_pass1")
        length, th, window = 10, 2, None
        tkw = dict(sigma_color_histogram=0.3, sigma_motion=3, alpha=0.7, cutoff_motion=0.1, cutoff_appearance=0.1,
                   max_distance=None, use_color_histogram=False)
        kind, wss2, mj2 = "tracklets_3d", 0.1, 30
    else:
        d = gu.load("c5_small_f30_p12")
        length, th, window = 5, 1, 3
        tkw = dict(sigma_color_histogram=0.3, sigma_motion=20, alpha=0.7, cutoff_motion=0.1, cutoff_appearance=0.1,
                   max_distance=60, use_color_histogram=case.endswith("_hist"))
        kind, wss2, mj2 = "tracklets_2d", 0.2, 12
    spec = _spec_from(d["spec"], mot3d)
    n = len(d["frame"])
    hist = None
    if case.endswith("_hist"):
        # _hist: every detection carries a histogram (a smooth pattern of its position + noise) and the tracklet cost
        # compares the medians over the 3-frame windows (k_window_median_hist vs DetectionTracklet2D's np.median)
        rng = np.random.RandomState(3)
        ph = np.concatenate([np.sin(d["pos"] / 40.0), np.cos(d["pos"] / 40.0)], 1)
        hist = (ph @ rng.randn(4, 12) + rng.randn(n, 12) * 0.4).reshape(n, 3, 4)
        d = dict(d, hist=hist)
    b = _batch_from(d, mot3d).to("cuda")
    gb, sol = mot3d.track_batch(b, spec)
    tracks1 = mot3d.tracks_to_lists(np.array([0, n]), sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(),
                                    sol.track_det.cpu().numpy())[0]
    tspec = mot3d.TrackletCostSpec(kind=kind, distance=tkw, confidence=dict(mul=1, bias=0), weight_source_sink=wss2,
                                   max_jump=mj2)
    got, info = mot3d.link_tracks(b, sol, tspec, split_length=length, th_length=th, window=window)
    # ---- the same through the objects
    three_d = kind == "tracklets_3d"
    Det = mot3d.Detection3D if three_d else mot3d.Detection2D
    dets = [Det(int(d["frame"][i]), d["pos"][i], confidence=float(d["conf"][i])) for i in range(n)]
    if hist is not None:
        for i, x in enumerate(dets):
            x.color_histogram = hist[i]
    index_of = {id(x): i for i, x in enumerate(dets)}
    pieces = []
    for t in tracks1:
        pieces += mot3d.split_trajectory_modulo([dets[i] for i in t], length=length)
    pieces = mot3d.remove_short_trajectories(pieces, th_length=th)
    Trk = mot3d.DetectionTracklet3D if three_d else mot3d.DetectionTracklet2D
    dts = [Trk(pc, window=window) for pc in pieces]
    assert info["n_tracklets"] == len(dts)
    fn = wf.weight_distance_tracklets_3d if three_d else wf.weight_distance_tracklets_2d
    g2 = mot3d.build_graph(dts, weight_source_sink=wss2, max_jump=mj2, verbose=False,
                           weight_confidence=lambda t: wf.weight_confidence_tracklets_2d(t, mul=1, bias=0),
                           weight_distance=lambda a, c: fn(a, c, **tkw))
    assert g2 is not None and g2.n_transitions == info["n_arcs"]
    final = mot3d.solve_graph(g2, verbose=False)
    want = [[index_of[id(x)] for x in mot3d.concat_tracklets([t.tracklet for t in tr])] for tr in final]
    assert g2.total_cost == info["total_cost"]
    assert sorted(got) == sorted(want)
    # the tracklets themselves: same pieces, ordered by the frame they start in
    first, ln, key = (x.cpu().numpy() for x in info["tracklets"])
    td = sol.track_det.cpu().numpy()
    mine = sorted(tuple(td[f:f + l].tolist()) for f, l in zip(first, ln))
    assert mine == sorted(tuple(index_of[id(x)] for x in pc) for pc in pieces)
    assert (np.diff(key) >= 0).all()


def test_c3_with_baseline_max_jump_4():
    """BASELINE.json quotes soldiers_3d with max_jump=4 (the example itself uses 8, which the golden vector pins): the
    same detections with max_jump=4 against the oracle - arcs, K, total, tracks - and against the live muSSP where it
    was compiled (total; its tolerances may leave it a few units above the exact optimum)."""
    import mot3d_b200 as mot3d
    from oracle import oracle
    d = gu.load("c3_soldiers_pass1")
    s = dict(d["spec"], max_jump=4)
    spec = _spec_from(s, mot3d)
    n = len(d["frame"])
    b = _batch_from(d, mot3d).to("cuda")
    gb, sol = mot3d.track_batch(b, spec, want_path_costs=True)
    ospec = oracle.CostSpec(**{k: v for k, v in s.items() if k != "kind"})
    row_ptr, col, w = oracle.build_transitions(d["frame"], d["pos"], ospec)
    assert gb.n_edges == len(col)
    tail, head, ww = oracle.graph_arcs(n, float(ospec.weight_source_sink), oracle.node_weights(d["conf"], ospec), 0.0,
                                       row_ptr, col, w)
    cost = oracle.quantise(ww)
    ref = oracle.ssp_solve(2 * n + 2, tail, head, cost)
    K = int(sol.n_paths[0].item())
    assert K == ref["K"] and int(sol.total_cost[0].item()) == ref["total"]
    assert (sol.path_cost[:K].cpu().numpy() == ref["path_cost"]).all()
    tracks = mot3d.tracks_to_lists(np.array([0, n]), sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(),
                                   sol.track_det.cpu().numpy())[0]
    assert tracks == oracle.tracks_from_flow(n, tail, head, ref["flow"])
    if oracle.have_ref():
        live = oracle.mussp_solve(2 * n + 2, tail, head, np.round(ww, 7))
        assert 0 <= oracle.flow_cost(cost, live["flow"]) - ref["total"] <= 2


@pytest.mark.parametrize("shape", ["crowd_big", "crowd_thin", "crowd_wide_costs", "band1", "band2"])
def test_large_components_against_oracle(shape):
    """Components beyond band 0 of the solver, one connected crowd each, against the integer oracle (K, total, every
    path cost, a feasible flow of that cost): solve_big.cuh (crowded: bucketed rounds over pushed work lists),
    solve_global.cuh (thin components, and costs beyond 32 bits), and the larger shared-memory bands of solve_sm.cuh."""
    import mot3d_b200 as mot3d
    from oracle import oracle
    P, F, extent, wss, seed = {"crowd_big": (64, 48, 140.0, 1.0, 11), "crowd_thin": (18, 170, 70.0, 1.0, 12),
                               "crowd_wide_costs": (40, 40, 110.0, 1.0, 13), "band1": (7, 45, 60.0, 1.0, 14),
                               "band2": (14, 55, 75.0, 0.5, 15)}[shape]
    rng = np.random.RandomState(seed)
    pos = rng.rand(P, 2) * extent
    vel = rng.randn(P, 2) * 1.2
    frames, poss, confs = [], [], []
    for f in range(F):
        pos = pos + vel + rng.randn(P, 2) * 0.5
        vel = np.where((pos < 0) | (pos > extent), -vel, vel)  # people stay in the room: one connected crowd
        keep = rng.rand(P) < 0.93
        frames.append(np.full(int(keep.sum()), f))
        poss.append(pos[keep] + rng.randn(int(keep.sum()), 2) * 0.8)
        confs.append(0.4 + 0.6 * rng.rand(int(keep.sum())))
    frame, xy, conf = np.concatenate(frames), np.concatenate(poss), np.concatenate(confs)
    n = len(frame)
    kw = dict(sigma_jump=1, sigma_distance=9, max_distance=28, use_color_histogram=False, use_bbox=False)
    mul = 400 if shape == "crowd_wide_costs" else 1  # node costs of -(160 .. 400): beyond 32 bits in 1e-7 units
    spec = mot3d.CostSpec(kind="detections_2d", distance=kw, confidence=dict(mul=mul, bias=0), weight_source_sink=wss,
                          max_jump=5)
    b, _ = mot3d.DetectionBatch.from_arrays(np.array([0, n]), frame, xy, conf)
    gb, sol = mot3d.track_batch(b.to("cuda"), spec, want_path_costs=True)
    ospec = oracle.CostSpec(weight_source_sink=wss, max_jump=5, mul=mul, bias=0, **kw)
    row_ptr, col, w = oracle.build_transitions(frame, xy, ospec)
    assert gb.n_edges == len(col)
    tail, head, ww = oracle.graph_arcs(n, float(wss), oracle.node_weights(conf, ospec), 0.0, row_ptr, col, w)
    cost = oracle.quantise(ww)
    ref = oracle.ssp_solve(2 * n + 2, tail, head, cost)
    K = int(sol.n_paths[0].item())
    assert K == ref["K"] and K >= 1
    assert int(sol.total_cost[0].item()) == ref["total"]
    assert (sol.path_cost[:K].cpu().numpy() == ref["path_cost"]).all()
    # the components really are large: the biggest one holds most of the detections
    stats = sol.stats.cpu().numpy()
    tracks = mot3d.tracks_to_lists(np.array([0, n]), sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(),
                                   sol.track_det.cpu().numpy())[0]
    arc = {(int(t), int(h)): int(c) for t, h, c in zip(tail, head, cost)}
    got, used = 0, set()
    for t in tracks:
        got += arc[(1, 2 + 2 * t[0])] + arc[(3 + 2 * t[-1], 2 * n + 2)]
        for a_, b_ in zip(t[:-1], t[1:]):
            got += arc[(3 + 2 * a_, 2 + 2 * b_)]
        for d_ in t:
            assert d_ not in used
            used.add(d_)
            got += arc[(2 + 2 * d_, 3 + 2 * d_)]
    assert got == ref["total"] and len(tracks) == K
    assert int(stats[11]) >= 1  # at least one branch was re-labelled: the solver proper ran


def test_solve_in_phases_equals_one_call():
    """mot3d_solve_batch_phase: BEGIN, FRONT over three ranges of graphs (as their rows become ready), BACK - the same
    flow, K and totals as mot3d_solve_batch in one call."""
    import ctypes
    import mot3d_b200 as mot3d
    from mot3d_b200 import _lib
    from mot3d_b200.batch import _ptr, _stream
    torch = _torch()
    d = gu.load("c5_small_f30_p12")
    spec = _spec_from(d["spec"], mot3d)
    n, G = len(d["frame"]), 7
    gp = np.arange(G + 1) * n
    b, _ = mot3d.DetectionBatch.from_arrays(gp, np.tile(d["frame"], G), np.tile(d["pos"], (G, 1)), np.tile(d["conf"], G))
    gb = mot3d.build_graphs(b.to("cuda"), spec)
    sol = mot3d.solve_graphs(gb)
    dev = gb.tkey.device
    L = _lib.lib()
    N = n * G
    nxt = torch.empty(N, dtype=torch.int32, device=dev)
    prv = torch.empty(N, dtype=torch.int32, device=dev)
    K = torch.zeros(G, dtype=torch.int32, device=dev)
    tot = torch.zeros(G, dtype=torch.int64, device=dev)
    wsb = L.mot3d_solve_workspace_bytes(N, G, gb.n_edges)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)

    def phase(ph, g0, gc):
        _lib.check(L.mot3d_solve_batch_phase(ph, g0, gc, G, N, _ptr(gb.dets.graph_ptr), n, _ptr(gb.tkey), gb.n_edges,
                                             _ptr(gb.row_ptr), _ptr(gb.col), _ptr(gb.cost), _ptr(gb.entry_cost),
                                             _ptr(gb.node_cost), _ptr(gb.exit_cost), _ptr(nxt), _ptr(prv), _ptr(K),
                                             _ptr(tot), None, None, _ptr(ws), wsb, _stream(dev)))
    phase(_lib.SOLVE_BEGIN, 0, 0)
    phase(_lib.SOLVE_FRONT, 0, 2)
    phase(_lib.SOLVE_FRONT, 2, 4)
    phase(_lib.SOLVE_FRONT | _lib.SOLVE_BACK, 6, 1)
    assert torch.equal(nxt, sol.next[:N]) and torch.equal(prv, sol.prev[:N])
    assert torch.equal(K, sol.n_paths[:G]) and torch.equal(tot, sol.total_cost[:G])
    assert tot.cpu().tolist() == [gu.ref_total_units(d)] * G
    with pytest.raises(mot3d.Mot3dError):
        phase(8, 0, 0)
