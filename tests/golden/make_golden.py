"""Generates tests/golden/*.npz from the reference itself. Run in the BUILD container only:

    python tests/golden/make_golden.py

It imports /root/reference/mot3d UNMODIFIED (imageio / matplotlib / the Boost wrapper are stubbed in
sys.modules: they are not installed here; the wrapper stub calls the real vendored muSSP through
oracle/_ref/libmussp_ref.so, fed the same text lines wrapper.solve would receive) and freezes, per vector:
inputs (sorted detections as arrays, cost parameters), the reference graph (arcs in graph_to_text order with the
float weights, their "{:0.7f}" quantisation, sha of the text) and the reference answer (K, path costs, total,
per-arc flow, trajectories as indices into the sorted detections).
The reference's own tests pin nothing; these files are the parity anchor.
/root/reference does not exist on the GPU box, hence the committed fixtures.
"""
import json
import os
import re
import sys
import types
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import oracle  # noqa: E402
from oracle import build_ref  # noqa: E402


def import_reference():
    assert build_ref.build(), "oracle/_ref could not be built"
    for name in ("imageio", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].use = lambda *a, **k: None
    wrapper = types.ModuleType("mot3d.solvers.wrappers.muSSP.wrapper")
    wrapper.solve = oracle.wrapper_solve
    sys.modules["mot3d.solvers.wrappers.muSSP.wrapper"] = wrapper
    sys.path.insert(0, REF)
    import mot3d  # the reference package
    import mot3d.solvers.wrappers.muSSP as pkg
    pkg.wrapper = wrapper
    return mot3d


def parse_text(lines):
    n, m = [int(x) for x in lines[0].split()[2:4]]
    tail, head, w, c = [], [], [], []
    for l in lines[1:]:
        if l[0] != "a":
            continue
        _, s, t, x = l.split()
        tail.append(int(s))
        head.append(int(t))
        w.append(float(x))
        neg = x.startswith("-")
        whole, frac = x.lstrip("-").split(".")
        v = int(whole) * 10 ** 7 + int(frac)
        c.append(-v if neg else v)
    return n, np.array(tail, np.int32), np.array(head, np.int32), np.array(w), np.array(c, np.int64)


def solve_reference(lines):
    n, tail, head, w, c = parse_text(lines)
    res = oracle.mussp_solve_text(lines, len(tail))
    return dict(n_nodes=n, tail=tail, head=head, w_text=w, cost=c, flow=res["flow"], K=res["K"],
                path_cost=res["path_cost"], total=res["total"], sha=oracle.text_sha(lines))


def freeze_detection_case(mot3d, name, detections, build_kwargs, spec, extra=None):
    """Run reference build_graph + solve_graph on Detection objects; store everything."""
    g = mot3d.build_graph(detections, verbose=False, **build_kwargs)
    assert g is not None
    lines = mot3d.graph_to_text(g)
    sol = solve_reference(lines)
    # exact float weights of the arcs (before text rounding), in g.edges() order
    w_float = np.array([d["weight"] for _, _, d in g.edges(data=True)], dtype=np.float64)
    trajectories = mot3d.solve_graph(g, verbose=False, method="muSSP")
    # sorted order = node order in g
    sorted_dets = [data["detection"] for _, data in g.nodes(data=True) if data.get("label") == "pre-node"]
    index_of = {id(d): i for i, d in enumerate(sorted_dets)}
    tracks = [[index_of[id(d)] for d in t] for t in trajectories]
    frame = np.array([d.index for d in sorted_dets], dtype=np.int64)
    pos = np.array([np.asarray(d.position, dtype=np.float64) for d in sorted_dets])
    conf = np.array([d.confidence for d in sorted_dets], dtype=np.float64)
    out = dict(sol)
    out.update(frame=frame, pos=pos, conf=conf, w_float=w_float,
               track_ptr=np.cumsum([0] + [len(t) for t in tracks]).astype(np.int64),
               track_det=np.array([i for t in tracks for i in t], dtype=np.int64),
               spec=json.dumps(spec))
    if extra:
        out.update(extra(sorted_dets))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-22s V=%d E=%d K=%d total=%.7f sha=%s (%d KB)" % (name, sol["n_nodes"], len(sol["tail"]), sol["K"],
                                                              sol["total"], sol["sha"], os.path.getsize(path) // 1024))
    return g, trajectories


def freeze_graph_case(name, lines, extra=None):
    sol = solve_reference(lines)
    if extra:
        sol.update(extra)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **sol)
    print("%-22s V=%d E=%d K=%d total=%.7f sha=%s (%d KB)" % (name, sol["n_nodes"], len(sol["tail"]), sol["K"],
                                                              sol["total"], sol["sha"], os.path.getsize(path) // 1024))



def tracklet_arrays(dts, dim):
    """Inputs of a tracklet graph as flat arrays: per tracklet the head / tail windows (indexes, positions), the
    median histograms of both ends (or absence), tail access flag, confidence."""
    hp, tp = [0], [0]
    hidx, tidx, hpos, tpos = [], [], [], []
    hh, th, has_h = [], [], []
    for t in dts:
        hidx += list(t.head.indexes)
        hpos += [np.asarray(p_, dtype=np.float64)[:dim] for p_ in t.head.positions]
        hp.append(len(hidx))
        tidx += list(t.tail.indexes)
        tpos += [np.asarray(p_, dtype=np.float64)[:dim] for p_ in t.tail.positions]
        tp.append(len(tidx))
        h1 = getattr(t.head, "color_histogram", None)
        h2 = getattr(t.tail, "color_histogram", None)
        has_h.append(h1 is not None and h2 is not None)
        hh.append(None if h1 is None else np.asarray(h1, dtype=np.float64))
        th.append(None if h2 is None else np.asarray(h2, dtype=np.float64))
    out = dict(t_head_ptr=np.array(hp), t_tail_ptr=np.array(tp), t_head_idx=np.array(hidx), t_tail_idx=np.array(tidx),
               t_head_pos=np.array(hpos), t_tail_pos=np.array(tpos),
               t_conf=np.array([t.confidence for t in dts], dtype=np.float64),
               t_tail_access=np.array([bool(getattr(t.tail, "access_point", False)) for t in dts]),
               t_has_hist=np.array(has_h))
    if any(has_h):
        shape = next(h.shape for h in hh if h is not None)
        out["t_head_hist"] = np.array([np.zeros(shape) if h is None else h for h in hh])
        out["t_tail_hist"] = np.array([np.zeros(shape) if h is None else h for h in th])
    if dts and hasattr(dts[0], "tracklets_2d"):
        # 3-D tracklets: per view (dict order) the median histograms of the 2-D tracklet's two ends
        names = sorted({v for t in dts for v in t.tracklets_2d})
        vid = {v: i for i, v in enumerate(names)}
        ptr, ids, vh, vt = [0], [], [], []
        for t in dts:
            for v, t2 in t.tracklets_2d.items():
                if t2.head.color_histogram is None or t2.tail.color_histogram is None:
                    continue
                ids.append(vid[v])
                vh.append(np.asarray(t2.head.color_histogram, dtype=np.float64))
                vt.append(np.asarray(t2.tail.color_histogram, dtype=np.float64))
            ptr.append(len(ids))
        if ids:
            out.update(t_view_ptr=np.array(ptr, np.int32), t_view_id=np.array(ids, np.int32),
                       t_view_head_hist=np.array(vh), t_view_tail_hist=np.array(vt))
    return out


def freeze_tracklet_case(mot3d, name, dts, build_kwargs, spec, dim):
    g = mot3d.build_graph(dts, verbose=False, **build_kwargs)
    assert g is not None
    lines = mot3d.graph_to_text(g)
    sol = solve_reference(lines)
    sol["w_float"] = np.array([d["weight"] for _, _, d in g.edges(data=True)], dtype=np.float64)
    trajectories = mot3d.solve_graph(g, verbose=False, method="muSSP")
    index_of = {id(d): i for i, d in enumerate(dts)}
    tracks = [[index_of[id(d)] for d in t] for t in trajectories]
    sol.update(track_ptr=np.cumsum([0] + [len(t) for t in tracks]).astype(np.int64),
               track_det=np.array([i for t in tracks for i in t], dtype=np.int64), spec=json.dumps(spec))
    sol.update(tracklet_arrays(dts, dim))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **sol)
    print("%-22s V=%d E=%d K=%d total=%.7f sha=%s (%d KB)" % (name, sol["n_nodes"], len(sol["tail"]), sol["K"],
                                                              sol["total"], sol["sha"], os.path.getsize(path) // 1024))
    return trajectories


def gen_synth(F, P, seed, extent=1000.0):
    """SURVEY.md 8(d) generator gen(F,P,seed)."""
    rng = np.random.RandomState(seed)
    pos = rng.rand(P, 2) * extent
    vel = rng.randn(P, 2) * 1.0
    frames, positions = [], []
    for f in range(F):
        pos = pos + vel + rng.randn(P, 2) * 0.2
        frames.append(np.full(P, f, dtype=np.int64))
        positions.append(pos.copy())
    return np.concatenate(frames), np.concatenate(positions)


def case_multiview(mot3d, wf):
    """(3b) synthetic 3-D detections seen in 2-4 of 4 camera views, both appearance terms of
    weight_distance_detections_3d on (mot3d/weight_functions.py:43-75: averages over the shared views)."""
    rng = np.random.RandomState(23)
    P, F, V, C, B = 6, 18, 4, 3, 24
    base = rng.rand(P, 3) * np.array([20.0, 20.0, 0.3]) + np.array([0, 0, 1.6])
    vel = rng.randn(P, 3) * np.array([0.15, 0.15, 0.0])
    proto = rng.rand(P, V, C, B)
    dets = []
    for f in range(F):
        base = base + vel + rng.randn(P, 3) * np.array([0.03, 0.03, 0.005])
        for p in range(P):
            if rng.rand() < 0.08:
                continue
            seen = [v for v in rng.permutation(V).tolist() if rng.rand() < 0.75]  # dict order differs per detection
            if len(seen) < 2:
                seen = [int(x) for x in rng.permutation(V)[:2]]
            d2 = {}
            for v in seen:
                h = proto[p, v] + 0.2 * rng.rand(C, B)
                h = h - h.mean(axis=1, keepdims=True)
                w, hh = 30 + rng.randn() * 3 + 4 * v, 80 + rng.randn() * 8 + 6 * v
                x, y = rng.rand(2) * 500
                d2["cam%d" % v] = mot3d.Detection2D(f, (x, y), view="cam%d" % v,
                                                    color_histogram=[h[c] for c in range(C)],
                                                    bbox=(x - w / 2, y - hh / 2, x + w / 2, y + hh / 2))
            dets.append(mot3d.Detection3D(f, base[p].copy(), confidence=0.6 + 0.4 * rng.rand(), detections_2d=d2))
    # views 0 and 1 are always shared here? no: make sure every linked pair shares a view (np.mean([]) is nan in the
    # reference): give every detection camera 0
    for d in dets:
        if "cam0" not in d.detections_2d:
            h = rng.rand(C, B)
            h = h - h.mean(axis=1, keepdims=True)
            d.detections_2d["cam0"] = mot3d.Detection2D(d.index, (0.0, 0.0), view="cam0",
                                                        color_histogram=[h[c] for c in range(C)],
                                                        bbox=(0.0, 0.0, 28.0 + rng.rand(), 75.0 + rng.rand()))
    spec = dict(kind="3d", sigma_jump=1, sigma_distance=0.6, sigma_color_histogram=0.35, sigma_box_size=0.25,
                max_distance=1.5, use_color_histogram=True, use_bbox=True, mul=1, bias=0.02,
                weight_source_sink=0.7, max_jump=3)
    wd = lambda d1, d2: wf.weight_distance_detections_3d(d1, d2, sigma_jump=1, sigma_distance=0.6,
                                                         sigma_color_histogram=0.35, sigma_box_size=0.25,
                                                         max_distance=1.5, use_color_histogram=True, use_bbox=True)
    wc = lambda d: wf.weight_confidence_detections_3d(d, mul=1, bias=0.02)

    def extra_views(sorted_dets):
        names = sorted({v for d in sorted_dets for v in d.detections_2d})
        vid = {v: i for i, v in enumerate(names)}
        ptr, ids, bb, hh = [0], [], [], []
        for d in sorted_dets:
            for v, d2 in d.detections_2d.items():  # dict order = the order the reference iterates in
                ids.append(vid[v])
                bb.append(d2.bbox)
                hh.append(np.asarray(d2.color_histogram, dtype=np.float64))
            ptr.append(len(ids))
        return dict(view_ptr=np.array(ptr, np.int32), view_id=np.array(ids, np.int32),
                    view_bbox=np.array(bb, np.float64), view_hist=np.array(hh, np.float64))

    freeze_detection_case(mot3d, "c3_synth_multiview", dets,
                          dict(weight_source_sink=0.7, max_jump=3, weight_confidence=wc, weight_distance=wd),
                          spec, extra_views)


ONLINE_CONFIG = dict(
    batch_size=50, completed_after=30,
    detections=dict(weight_source_sink=1, max_jump=4, verbose=False,
                    weight_distance=dict(sigma_distance=10, sigma_jump=4, sigma_color_histogram=0.1,
                                         sigma_box_size=0.1, max_distance=30, use_color_histogram=False,
                                         use_bbox=True),
                    weight_confidence=dict(mul=0, bias=0)),
    tracklets=dict(weight_source_sink=0.1, max_jump=40, length_endpoints=5, verbose=False,
                   weight_distance=dict(sigma_color_histogram=0.3, sigma_motion=100, alpha=0.2, cutoff_motion=0.1,
                                        cutoff_appearance=0.1, use_color_histogram=True),
                   weight_confidence=dict(mul=0, bias=0.11)))


def gen_online(F=200, P=8, seed=0):
    """C4-synth: P people over a 768 x 576 image, per-frame miss probability 0.1, Poisson(1) false
    positives, bbox (w ~ N(30,3), h ~ N(80,8)) around the position. Returns per-frame arrays."""
    rng = np.random.RandomState(seed)
    pos = rng.rand(P, 2) * np.array([768.0, 576.0])
    vel = rng.randn(P, 2) * 1.0
    frames = []
    for f in range(F):
        pos = pos + vel + rng.randn(P, 2) * 0.2
        rows = []
        for p in range(P):
            if rng.rand() < 0.1:
                continue
            if (p == 0 and f > 70) or (p == 1 and f > 95) or (p == 2 and f < 60):
                continue  # two people leave, one enters late: trajectories complete / start between batches
            rows.append((pos[p, 0], pos[p, 1], 0.75 + 0.25 * rng.rand()))
        for _ in range(rng.poisson(1.0)):
            rows.append((rng.rand() * 768, rng.rand() * 576, 0.75 + 0.25 * rng.rand()))
        out = []
        for x, y, c in rows:
            w, h = 30 + rng.randn() * 3, 80 + rng.randn() * 8
            out.append((x, y, c, x - w / 2, y - h / 2, x + w / 2, y + h / 2))
        frames.append(np.array(out, dtype=np.float64).reshape(-1, 7))
    return frames


def case_online(mot3d, with_hist=False):
    """(4d) the reference's own online loop (projects/PETS2009S2L1/singleview_tracking.py:76-134 Tracker2D.update,
    tracking_routines.py, scene.py) on C4-synth, 4 batches of 50 frames; the scene after every batch is frozen.
    with_hist: every detection also carries a colour histogram [3, 8] (a per-person pattern + noise; false positives
    random) and both costs use their appearance terms - c4_online_loop_hist."""
    import collections
    import collections.abc
    collections.Iterable = collections.abc.Iterable  # tracking_routines.py:9 predates Python 3.10
    sys.modules["multiview_matching"] = types.ModuleType("multiview_matching")  # not vendored, unused on this path
    sys.modules["imageio"].imread = lambda *a, **k: None
    sys.path.insert(0, os.path.join(REF, "projects/PETS2009S2L1"))
    import singleview_tracking as sv
    frames = gen_online()
    cfg = ONLINE_CONFIG
    hists = None
    if with_hist:
        import copy
        cfg = copy.deepcopy(ONLINE_CONFIG)
        cfg["detections"]["weight_distance"].update(use_color_histogram=True, sigma_color_histogram=0.5)
        cfg["tracklets"]["weight_distance"].update(alpha=0.5, sigma_color_histogram=0.4)
        # who made each detection is not recorded by gen_online: give nearby detections alike patterns instead - a
        # smooth random field of the position (people move ~1 px per frame) plus noise
        rng = np.random.RandomState(11)
        W = rng.randn(4, 24)
        hists = []
        for f, fr in enumerate(frames):
            ph = np.stack([np.sin(fr[:, 0] / 60.0), np.cos(fr[:, 0] / 60.0), np.sin(fr[:, 1] / 60.0),
                           np.cos(fr[:, 1] / 60.0)], 1) if len(fr) else np.zeros((0, 4))
            # from frame 120 on everybody in the left half of the image looks different (the pattern flips): links
            # that motion alone would make across that frame are cut by the appearance terms
            sign = np.where((f >= 120) & (fr[:, 0] < 384), -1.0, 1.0)[:, None] if len(fr) else np.zeros((0, 1))
            h = sign * (ph @ W) + rng.randn(len(fr), 24) * 0.25
            h = h.reshape(-1, 3, 8)
            hists.append(h - h.mean(axis=2, keepdims=True))
    scene = sv.Scene(cfg["completed_after"])
    tracker = sv.Tracker2D(scene, None, cfg, image_size=(768, 576))
    out = dict(det_ptr=np.cumsum([0] + [len(f) for f in frames]).astype(np.int64),
               det=np.concatenate(frames), config=json.dumps(cfg))
    B = cfg["batch_size"]
    import io
    import contextlib
    for b in range(len(frames) // B):
        dets = [[mot3d.Detection2D(f, r[:2].copy(), confidence=float(r[2]), bbox=tuple(r[3:7]),
                                   color_histogram=None if hists is None else hists[f][k].copy())
                 for k, r in enumerate(frames[f])] for f in range(b * B, (b + 1) * B)]
        with contextlib.redirect_stdout(io.StringIO()):
            tracker.update((b + 1) * B, dets)  # iterator.idx after the batch (feeders.py)
        for name, store in (("active", scene.active), ("completed", scene.completed)):
            ids = sorted(store)
            out["b%d_%s_ids" % (b, name)] = np.array(ids, dtype=np.int64)
            out["b%d_%s_ptr" % (b, name)] = np.cumsum([0] + [len(store[i]) for i in ids]).astype(np.int64)
            out["b%d_%s_index" % (b, name)] = np.array([d.index for i in ids for d in store[i]], dtype=np.int64)
            out["b%d_%s_pos" % (b, name)] = np.array([np.asarray(d.position, dtype=np.float64) for i in ids
                                                      for d in store[i]]).reshape(-1, 2)
            out["b%d_%s_bbox" % (b, name)] = np.array([np.asarray(d.bbox, dtype=np.float64) for i in ids
                                                       for d in store[i]]).reshape(-1, 4)
        print("online batch %d: active %d completed %d tracklets %d" % (b, len(scene.active), len(scene.completed),
                                                                        len(scene.tracklets)))
    out["n_tracklets"] = len(scene.tracklets)
    name = "c4_online_loop_hist" if with_hist else "c4_online_loop"
    if with_hist:
        out["det_hist"] = np.concatenate(hists)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-22s (%d KB)" % (name, os.path.getsize(path) // 1024))


def case_multiview_tracklets(mot3d, wf):
    """(3c) tracklet linking of 3-D tracklets whose detections carry 2-D views with histograms:
    weight_distance_tracklets_3d with use_color_histogram=True (mot3d/weight_functions.py:125-160: appearance averaged
    over the shared views of the per-view 2-D tracklets, mot3d/types.py:142-184)."""
    rng = np.random.RandomState(31)
    P, F, V, C, B = 5, 60, 3, 3, 16
    base = rng.rand(P, 3) * np.array([20.0, 20.0, 0.3]) + np.array([0, 0, 1.6])
    vel = rng.randn(P, 3) * np.array([0.12, 0.12, 0.0])
    proto = rng.rand(P, V, C, B)
    tracks = [[] for _ in range(P)]
    for f in range(F):
        base = base + vel + rng.randn(P, 3) * np.array([0.02, 0.02, 0.004])
        for p in range(P):
            if (f // 6) % 3 == 2 and p % 2 == 0:
                continue  # gaps that the linking has to bridge
            d2 = {}
            for v in [0] + [v for v in (1, 2) if rng.rand() < 0.7]:
                h = proto[p, v] + 0.2 * rng.rand(C, B)
                h = h - h.mean(axis=1, keepdims=True)
                d2["cam%d" % v] = mot3d.Detection2D(f, tuple(rng.rand(2) * 500), view="cam%d" % v,
                                                    color_histogram=[h[c] for c in range(C)], bbox=(0, 0, 30, 80))
            tracks[p].append(mot3d.Detection3D(f, base[p].copy(), detections_2d=d2))
    dts = []
    from mot3d.utils import trajectory as ref_traj
    for tr in tracks:
        for piece in ref_traj.split_trajectory_modulo(tr, length=6):
            if len(piece) > 2:
                dts.append(mot3d.DetectionTracklet3D(piece, window=4))
    order = rng.permutation(len(dts))
    dts = [dts[i] for i in order]
    spec = dict(kind="tracklets_3d", sigma_color_histogram=0.3, sigma_motion=1.5, alpha=0.6, cutoff_motion=0.1,
                cutoff_appearance=0.1, max_distance=6.0, use_color_histogram=True, mul=0, bias=0.11,
                weight_source_sink=0.1, max_jump=14)
    wd = lambda t1, t2: wf.weight_distance_tracklets_3d(t1, t2, sigma_color_histogram=0.3, sigma_motion=1.5,
                                                        alpha=0.6, cutoff_motion=0.1, cutoff_appearance=0.1,
                                                        max_distance=6.0, use_color_histogram=True)
    wc = lambda t: wf.weight_confidence_tracklets_3d(t, mul=0, bias=0.11)
    freeze_tracklet_case(mot3d, "c3_synth_multiview_tracklets", dts,
                         dict(weight_source_sink=0.1, max_jump=14, weight_confidence=wc, weight_distance=wd), spec, 3)


def case_merge_close_points(mot3d):
    """(3a) the de-duplication in front of the 3-D example (examples/soldiers_3d/example.py:20-35): the reference's
    merge_close_points (mot3d/utils/filt.py:6-24, sklearn DBSCAN) on every frame of detections_3d_head.json used
    there, eps 0.25; plus return_indexes on a crowded synthetic set."""
    from mot3d.utils import utils as ref_utils
    from mot3d.utils import filt as ref_filt
    det = ref_utils.json_read(os.path.join(REF, "examples/soldiers_3d/detections_3d_head.json"))
    raw, merged = [], []
    for index_frame in range(49650, 50000):
        pts = det[str(index_frame)]
        raw.append(np.asarray(pts, dtype=np.float64).reshape(-1, 3))
        merged.append(np.asarray(ref_filt.merge_close_points(pts, eps=0.25), dtype=np.float64).reshape(-1, 3))
    rng = np.random.RandomState(5)
    crowd = rng.rand(400, 2) * 10.0
    groups = ref_filt.merge_close_points(crowd, eps=0.3, return_indexes=True)
    crowd_merged = ref_filt.merge_close_points(crowd, eps=0.3)
    out = dict(raw_ptr=np.cumsum([0] + [len(r) for r in raw]).astype(np.int64), raw=np.concatenate(raw),
               merged_ptr=np.cumsum([0] + [len(r) for r in merged]).astype(np.int64), merged=np.concatenate(merged),
               crowd=crowd, crowd_merged=np.asarray(crowd_merged),
               crowd_group_ptr=np.cumsum([0] + [len(g) for g in groups]).astype(np.int64),
               crowd_group_idx=np.array([i for g in groups for i in g], dtype=np.int64))
    path = os.path.join(HERE, "c3_merge_close_points.npz")
    np.savez_compressed(path, **out)
    print("%-22s frames=%d points %d -> %d; crowd 400 -> %d (%d KB)" % ("c3_merge_close_points", len(raw),
                                                                         len(out["raw"]), len(out["merged"]),
                                                                         len(groups), os.path.getsize(path) // 1024))


def case_color_histogram(mot3d):
    """(N3) the reference's color_histogram (mot3d/distance_functions.py:10-19, cv2 + numpy) on patches of a seeded
    synthetic RGB frame, cropped the way the feeder does (projects/PETS2009S2L1/feeders.py:104:
    img[ymin:ymax, xmin:xmax]): default arguments, a channel subset, other bin counts, boxes reaching beyond the
    image (numpy clips the slice)."""
    rng = np.random.RandomState(11)
    H, W = 240, 320
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.stack([(xx * 255 // (W - 1)), (yy * 255 // (H - 1)), ((xx + yy) * 255 // (H + W - 2))], axis=-1)
    img = np.clip(img + rng.randint(-40, 41, size=img.shape), 0, 255).astype(np.uint8)
    img[100:140, 200:260] = rng.randint(0, 256, size=(40, 60, 3))  # a patch of arbitrary colours
    img[0:30, 0:30] = 255
    img[30:60, 0:30] = 0
    boxes = [(10, 20, 60, 140), (200, 100, 260, 140), (0, 0, 30, 60), (150, 30, 151, 31), (300, 200, 340, 260),
             (5, 5, 320, 240), (100, 50, 180, 230), (250, 10, 319, 90)]
    for _ in range(8):
        x0, y0 = rng.randint(0, W - 8), rng.randint(0, H - 8)
        boxes.append((x0, y0, x0 + rng.randint(4, 90), y0 + rng.randint(4, 120)))
    boxes = np.array(boxes, dtype=np.int32)
    out = dict(image=img, boxes=boxes)
    for tag, kw in (("default", {}), ("ch12_n64", dict(channels=(1, 2), N=64)), ("ch0_n100", dict(channels=(0,), N=100)),
                    ("ch210_n300", dict(channels=(2, 1, 0), N=300))):
        hs = []
        for xmin, ymin, xmax, ymax in boxes:
            hs.append(np.array(mot3d.color_histogram(img[ymin:ymax, xmin:xmax], **kw), dtype=np.float64))
        out["hist_" + tag] = np.stack(hs)
    path = os.path.join(HERE, "color_histogram.npz")
    np.savez_compressed(path, **out)
    print("%-22s boxes=%d (%d KB)" % ("color_histogram", len(boxes), os.path.getsize(path) // 1024))


def case_clipped_arcs(mot3d, wf):
    """(a9) a graph on which muSSP's Graph::invalid_edge_rm (Graph.cpp:86-111) really clips arcs: entering costs
    -0.3 (weight_source_sink < 0), so every transition weaker than that (w > exit + entry = -0.3) is removed by the
    reference before it solves. Noisy detections with drop-outs make about a third of the arcs that weak."""
    rng = np.random.RandomState(9)
    F, P = 30, 10
    pos = rng.rand(P, 2) * 100.0
    vel = rng.randn(P, 2) * 1.5
    dets = []
    for f in range(F):
        pos = pos + vel + rng.randn(P, 2) * 1.0
        for q in range(P):
            if rng.rand() < 0.85:
                dets.append(mot3d.Detection2D(f, pos[q] + rng.randn(2) * 1.5, confidence=float(0.3 + 0.6 * rng.rand())))
    kw = dict(sigma_jump=1, sigma_distance=6, max_distance=25, use_color_histogram=False, use_bbox=False)
    spec = dict(kind="2d", sigma_color_histogram=0.3, sigma_box_size=0.3, mul=1, bias=0, weight_source_sink=-0.3,
                max_jump=4, **kw)
    wd = lambda d1, d2: wf.weight_distance_detections_2d(d1, d2, **kw)
    wc = lambda d: wf.weight_confidence_detections_2d(d, mul=1, bias=0)

    def extra(sorted_dets):
        return {}
    g, _ = freeze_detection_case(mot3d, "a9_clipped_arcs", dets,
                                 dict(weight_source_sink=-0.3, max_jump=4, weight_confidence=wc, weight_distance=wd),
                                 spec, extra)
    w = np.array([d["weight"] for u, v, d in g.edges(data=True) if u % 2 == 1 and u != 1 and v != len(g)])
    print("    transitions %d, clipped by invalid_edge_rm (w > -0.3): %d" % (len(w), int((w > -0.3).sum())))


def main():
    mot3d = import_reference()
    import mot3d.weight_functions as wf
    if "a9_clipped_arcs" in sys.argv[1:]:
        return case_clipped_arcs(mot3d, wf)
    if "color_histogram" in sys.argv[1:]:
        return case_color_histogram(mot3d)
    if "c3_merge_close_points" in sys.argv[1:]:
        return case_merge_close_points(mot3d)
    if "c3_synth_multiview_tracklets" in sys.argv[1:]:
        return case_multiview_tracklets(mot3d, wf)
    if "c4_online_loop" in sys.argv[1:]:
        return case_online(mot3d)
    if "c4_online_loop_hist" in sys.argv[1:]:
        return case_online(mot3d, with_hist=True)
    if "c3_synth_multiview" in sys.argv[1:]:
        return case_multiview(mot3d, wf)
    case_color_histogram(mot3d)
    case_clipped_arcs(mot3d, wf)
    case_merge_close_points(mot3d)
    case_multiview(mot3d, wf)
    case_multiview_tracklets(mot3d, wf)
    case_online(mot3d)
    case_online(mot3d, with_hist=True)
    from mot3d.utils import utils as ref_utils
    from mot3d.utils import trajectory as ref_traj

    # ---- (0) the wrapper's own smoke graph (mot3d/solvers/wrappers/muSSP/test.py:3-71)
    src = open(os.path.join(REF, "mot3d/solvers/wrappers/muSSP/test.py")).read()
    lines = [l for l in re.findall(r'"([^"]+)"', src) if l[0] in "pac"]
    freeze_graph_case("wrapper_test_graph", lines)

    # ---- (1) C1 simple_2d (examples/simple_2d/example.py:17-58), jitter seeded
    ex = os.path.join(REF, "examples/simple_2d")
    ps1 = np.array(ref_utils.json_read(os.path.join(ex, "track_line1.json")))
    ps2 = np.array(ref_utils.json_read(os.path.join(ex, "track_line2.json")))
    ps2 = ps2 - np.array([[4, 0]])
    rng = np.random.RandomState(0)
    ps1[4] = ps1[4] + rng.randn(2)
    ps2[4] = ps2[4] + rng.randn(2)
    dummy_hist = [np.ones((128,))] * 3
    dummy_bbox = [0, 0, 1, 1]
    dets = []
    for index, pair in enumerate(zip(ps1, ps2)):
        for position in pair:
            dets.append(mot3d.Detection2D(index, position, color_histogram=dummy_hist, bbox=dummy_bbox))
    spec = dict(kind="2d", sigma_jump=1, sigma_distance=10, sigma_color_histogram=0.3, sigma_box_size=0.3,
                max_distance=15, use_color_histogram=True, use_bbox=True, mul=1, bias=0,
                weight_source_sink=0.1, max_jump=4)
    wd = lambda d1, d2: wf.weight_distance_detections_2d(d1, d2, sigma_jump=1, sigma_distance=10,
                                                         sigma_color_histogram=0.3, sigma_box_size=0.3,
                                                         max_distance=15, use_color_histogram=True, use_bbox=True)
    wc = lambda d: wf.weight_confidence_detections_2d(d, mul=1, bias=0)

    def extra_2d(sorted_dets):
        return dict(bbox=np.array([d.bbox for d in sorted_dets], dtype=np.float64),
                    hist=np.array([np.asarray(d.color_histogram, dtype=np.float64) for d in sorted_dets]))

    g, trajectories = freeze_detection_case(mot3d, "c1_simple_2d", dets,
                                            dict(weight_source_sink=0.1, max_jump=4, weight_confidence=wc,
                                                 weight_distance=wd), spec, extra_2d)

    # ---- (2) C2 simple_2d_tracklets pass 2 (examples/simple_2d_tracklets/example.py:78-108)
    tracklets = []
    for traj in trajectories:
        tracklets += mot3d.split_trajectory_modulo(traj, length=5)
    tracklets = mot3d.remove_short_trajectories(tracklets, th_length=2)
    dts = [mot3d.DetectionTracklet2D(t) for t in tracklets]
    wdt = lambda t1, t2: wf.weight_distance_tracklets_2d(t1, t2, max_distance=None, sigma_color_histogram=0.3,
                                                         sigma_motion=50, alpha=0.7, cutoff_motion=0.1,
                                                         cutoff_appearance=0.1, use_color_histogram=True)
    wct = lambda t: wf.weight_confidence_tracklets_2d(t, mul=1, bias=0)
    tspec = dict(kind="tracklets_2d", sigma_color_histogram=0.3, sigma_motion=50, alpha=0.7, cutoff_motion=0.1,
                 cutoff_appearance=0.1, max_distance=None, use_color_histogram=True, mul=1, bias=0,
                 weight_source_sink=0.1, max_jump=20)
    freeze_tracklet_case(mot3d, "c2_tracklets_pass2", dts,
                         dict(weight_source_sink=0.1, max_jump=20, weight_confidence=wct, weight_distance=wdt), tspec, 2)

    # ---- (3) C3 soldiers_3d (examples/soldiers_3d/example.py:25-101)
    ex = os.path.join(REF, "examples/soldiers_3d")
    det3d = ref_utils.json_read(os.path.join(ex, "detections_3d_head.json"))
    from mot3d.utils import filt as ref_filt
    dets = []
    for index, index_frame in enumerate(range(49650, 50000)):
        positions = ref_filt.merge_close_points(det3d[str(index_frame)], eps=0.25)
        for position in positions:
            dets.append(mot3d.Detection3D(index, position))
    spec = dict(kind="3d", sigma_jump=2, sigma_distance=0.15, sigma_color_histogram=0.3, sigma_box_size=0.3,
                max_distance=0.4, use_color_histogram=False, use_bbox=False, mul=1, bias=0,
                weight_source_sink=50, max_jump=8)
    wd = lambda d1, d2: wf.weight_distance_detections_3d(d1, d2, sigma_jump=2, sigma_distance=0.15,
                                                         sigma_color_histogram=0.3, sigma_box_size=0.3,
                                                         max_distance=0.4, use_color_histogram=False,
                                                         use_bbox=False)
    wc = lambda d: wf.weight_confidence_detections_3d(d, mul=1, bias=0)
    g, trajectories = freeze_detection_case(mot3d, "c3_soldiers_pass1", dets,
                                            dict(weight_source_sink=50, max_jump=8, weight_confidence=wc,
                                                 weight_distance=wd), spec)
    tracklets = []
    for traj in trajectories:
        tracklets += mot3d.split_trajectory_modulo(traj, length=10)
    tracklets = mot3d.remove_short_trajectories(tracklets, th_length=2)
    dts = [mot3d.DetectionTracklet3D(t) for t in tracklets]
    wdt = lambda t1, t2: wf.weight_distance_tracklets_3d(t1, t2, sigma_color_histogram=0.3, sigma_motion=3,
                                                         alpha=0.7, cutoff_motion=0.1, cutoff_appearance=0.1,
                                                         max_distance=None, use_color_histogram=False)
    wct = lambda t: wf.weight_confidence_tracklets_3d(t, mul=1, bias=0)
    tspec = dict(kind="tracklets_3d", sigma_color_histogram=0.3, sigma_motion=3, alpha=0.7, cutoff_motion=0.1,
                 cutoff_appearance=0.1, max_distance=None, use_color_histogram=False, mul=1, bias=0,
                 weight_source_sink=0.1, max_jump=30)
    freeze_tracklet_case(mot3d, "c3_soldiers_pass2", dts,
                         dict(weight_source_sink=0.1, max_jump=30, weight_confidence=wct, weight_distance=wdt), tspec, 3)

    # ---- (4) C4-real: PETS View_001 detections JSON, two 50-frame windows, bbox term on, no histograms
    #      (feeder logic of projects/PETS2009S2L1/feeders.py:51-124; parameters of
    #       config/config_singleview_PETS2009S2L1_View_001.yaml:24-38 with use_color_histogram=False)
    pets = ref_utils.json_read(os.path.join(REF, "projects/PETS2009S2L1/detections",
                                            "keypointsrcnn_resnet50_fpn_View_001_finetuned_2.json"))
    spec = dict(kind="2d", sigma_jump=4, sigma_distance=10, sigma_color_histogram=0.1, sigma_box_size=0.1,
                max_distance=30, use_color_histogram=False, use_bbox=True, mul=0, bias=0,
                weight_source_sink=1, max_jump=4)
    wd = lambda d1, d2: wf.weight_distance_detections_2d(d1, d2, sigma_distance=10, sigma_jump=4,
                                                         sigma_color_histogram=0.1, sigma_box_size=0.1,
                                                         max_distance=30, use_color_histogram=False, use_bbox=True)
    wc = lambda d: wf.weight_confidence_detections_2d(d, mul=0, bias=0)

    def extra_bbox(sorted_dets):
        return dict(bbox=np.array([d.bbox for d in sorted_dets], dtype=np.float64))

    def pets_window(f0, f1):
        dets = []
        for idx in range(f0, f1):
            boxes = [[max(int(x), 0) for x in box] for box in pets["bboxes"][idx]]
            kpss = [[p[:2] for p in kps] for kps in pets["keypoints"][idx]]
            for box, kps, score, label in zip(boxes, kpss, pets["scores"][idx], pets["labels"][idx]):
                if label == 1 and score > 0.75:
                    position = np.mean(np.reshape(kps, (-1, 2)), axis=0)[:2]
                    dets.append(mot3d.Detection2D(idx, position, confidence=score, bbox=tuple(box)))
        return dets

    for k, (f0, f1) in enumerate([(0, 50), (50, 100)]):
        freeze_detection_case(mot3d, "c4_pets_view1_w%d" % k, pets_window(f0, f1),
                              dict(weight_source_sink=1, max_jump=4, weight_confidence=wc, weight_distance=wd),
                              spec, extra_bbox)

    # ---- (4b) synthetic appearance case: random mean-subtracted histograms (3 x 32 bins) + bboxes, both terms on
    rng = np.random.RandomState(7)
    dets = []
    P, F = 6, 20
    base = rng.rand(P, 2) * 200
    vel = rng.randn(P, 2) * 2
    proto = rng.rand(P, 3, 32)
    for f in range(F):
        base = base + vel + rng.randn(P, 2) * 0.5
        for p in range(P):
            if rng.rand() < 0.1:
                continue
            h = proto[p] + 0.15 * rng.rand(3, 32)
            h = h - h.mean(axis=1, keepdims=True)
            w, hh = 30 + rng.randn() * 3, 80 + rng.randn() * 8
            x, y = base[p]
            dets.append(mot3d.Detection2D(f, base[p].copy(), confidence=0.75 + 0.25 * rng.rand(),
                                          color_histogram=[h[0], h[1], h[2]],
                                          bbox=(x - w / 2, y - hh / 2, x + w / 2, y + hh / 2)))
    spec = dict(kind="2d", sigma_jump=1, sigma_distance=10, sigma_color_histogram=0.3, sigma_box_size=0.3,
                max_distance=30, use_color_histogram=True, use_bbox=True, mul=1, bias=0.05,
                weight_source_sink=0.8, max_jump=3)
    wd = lambda d1, d2: wf.weight_distance_detections_2d(d1, d2, sigma_jump=1, sigma_distance=10,
                                                         sigma_color_histogram=0.3, sigma_box_size=0.3,
                                                         max_distance=30, use_color_histogram=True, use_bbox=True)
    wc = lambda d: wf.weight_confidence_detections_2d(d, mul=1, bias=0.05)
    freeze_detection_case(mot3d, "c4_synth_appearance", dets,
                          dict(weight_source_sink=0.8, max_jump=3, weight_confidence=wc, weight_distance=wd),
                          spec, extra_2d)

    # ---- (4c) synthetic tracklet linking with everything on: end windows, histograms on some tracklets only,
    #      access points, max_distance gate (PETS pass-2 shaped: config_singleview_*.yaml:40-55)
    rng = np.random.RandomState(11)
    dts = []
    P = 7
    base = rng.rand(P, 2) * 300
    vel = rng.randn(P, 2) * 1.5
    proto = rng.rand(P, 3, 16)
    for p_ in range(P):
        f = int(rng.randint(0, 6))
        pos = base[p_].copy()
        while f < 70:
            ln = int(rng.randint(3, 9))
            chain = []
            with_hist = rng.rand() < 0.8
            for k in range(ln):
                pos = pos + vel[p_] + rng.randn(2) * 0.4
                h = None
                if with_hist:
                    hh_ = proto[p_] + 0.2 * rng.rand(3, 16)
                    hh_ = hh_ - hh_.mean(axis=1, keepdims=True)
                    h = [hh_[0], hh_[1], hh_[2]]
                chain.append(mot3d.Detection2D(f + k, pos.copy(), confidence=0.9, color_histogram=h,
                                               bbox=(pos[0] - 15, pos[1] - 40, pos[0] + 15, pos[1] + 40)))
            dts.append(mot3d.DetectionTracklet2D(chain, window=4, confidence=0.5 + 0.5 * rng.rand(),
                                                 tail_access_point=bool(rng.rand() < 0.1)))
            gap = int(rng.randint(1, 6))
            pos = pos + vel[p_] * gap
            f += ln + gap
    order = rng.permutation(len(dts))
    dts = [dts[i] for i in order]
    tspec = dict(kind="tracklets_2d", sigma_color_histogram=0.3, sigma_motion=20, alpha=0.2, cutoff_motion=0.1,
                 cutoff_appearance=0.1, max_distance=60, use_color_histogram=True, mul=0, bias=0.11,
                 weight_source_sink=0.1, max_jump=12)
    wdt = lambda t1, t2: wf.weight_distance_tracklets_2d(t1, t2, sigma_color_histogram=0.3, sigma_motion=20, alpha=0.2,
                                                         cutoff_motion=0.1, cutoff_appearance=0.1, max_distance=60,
                                                         use_color_histogram=True)
    wct = lambda t: wf.weight_confidence_tracklets_2d(t, mul=0, bias=0.11)
    freeze_tracklet_case(mot3d, "c4_synth_tracklets", dts,
                         dict(weight_source_sink=0.1, max_jump=12, weight_confidence=wct, weight_distance=wdt), tspec, 2)

    # ---- (5) C5 window: gen(F=50, P=100, seed 0) = one 50-frame window of the batched config
    spec = dict(kind="2d", sigma_jump=1, sigma_distance=10, sigma_color_histogram=0.3, sigma_box_size=0.3,
                max_distance=30, use_color_histogram=False, use_bbox=False, mul=1, bias=0,
                weight_source_sink=1, max_jump=8)
    wd = lambda d1, d2: wf.weight_distance_detections_2d(d1, d2, sigma_jump=1, sigma_distance=10,
                                                         max_distance=30, use_color_histogram=False, use_bbox=False)
    wc = lambda d: wf.weight_confidence_detections_2d(d, mul=1, bias=0)
    for nm, F, P in (("c5_window_f50_p100", 50, 100), ("c5_small_f30_p12", 30, 12)):
        frame, pos = gen_synth(F, P, 0, extent=1000.0 if P == 100 else 120.0)
        dets = [mot3d.Detection2D(int(f), p, confidence=0.9) for f, p in zip(frame, pos)]
        freeze_detection_case(mot3d, nm, dets,
                              dict(weight_source_sink=1, max_jump=8, weight_confidence=wc, weight_distance=wd),
                              spec)

    # ---- (6) quirk: lone detection whose only path costs +0.5 is still returned (wrapper.cpp:205,220)
    dets = [mot3d.Detection2D(0, np.array([1.0, 2.0]), confidence=0.5)]
    spec = dict(kind="2d", sigma_jump=1, sigma_distance=2, sigma_color_histogram=0.3, sigma_box_size=0.3,
                max_distance=20, use_color_histogram=False, use_bbox=False, mul=1, bias=0,
                weight_source_sink=1.0, max_jump=4)
    wd = lambda d1, d2: wf.weight_distance_detections_2d(d1, d2, use_color_histogram=False, use_bbox=False)
    freeze_detection_case(mot3d, "quirk_single_detection", dets,
                          dict(weight_source_sink=1.0, max_jump=4, weight_confidence=wc, weight_distance=wd), spec)

    # ---- (7) muSSP's own sample graph (zip!/muSSP-master/input_MOT_seq07_followme.txt): solver-only vector
    with zipfile.ZipFile(os.path.join(REF, "mussp-master-6cf61b8.zip")) as z:
        txt = z.read("muSSP-master/input_MOT_seq07_followme.txt").decode()
    lines = [l for l in txt.split("\n") if l and l[0] in "pa"]
    n, tail, head, w, c = parse_text(lines)
    res = oracle.mussp_solve(n, tail, head, w)
    # the file carries 6 decimals at most? keep exact ints at 1e-7 via decimal parsing of the text
    out = dict(n_nodes=n, tail=tail, head=head, cost=c, K=res["K"], total=res["total"], path_cost=res["path_cost"],
               flow=res["flow"].astype(np.int8))
    path = os.path.join(HERE, "mussp_seq07.npz")
    np.savez_compressed(path, **out)
    print("%-22s V=%d E=%d K=%d total=%.7f (%d KB)" % ("mussp_seq07", n, len(tail), res["K"], res["total"],
                                                       os.path.getsize(path) // 1024))

    # ---- (8) C5 whole (F=1000, P=100): arcs from the numpy oracle (validated above on the window), real muSSP
    frame, pos = gen_synth(1000, 100, 0)
    ospec = oracle.CostSpec(sigma_jump=1, sigma_distance=10, max_distance=30, mul=1, bias=0, weight_source_sink=1,
                            max_jump=8)
    row_ptr, col, w = oracle.build_transitions(frame, pos, ospec)
    nw = oracle.node_weights(np.full(frame.shape[0], 0.9), ospec)
    tail, head, ww = oracle.graph_arcs(frame.shape[0], 1.0, nw, 0.0, row_ptr, col, w)
    c = oracle.quantise(ww)
    res = oracle.mussp_solve(2 * frame.shape[0] + 2, tail, head, c.astype(np.float64) * 1e-7)
    out = dict(n_nodes=2 * frame.shape[0] + 2, n_arcs=len(tail), n_transitions=len(col), K=res["K"],
               total=res["total"], path_cost=res["path_cost"], cost_sum_check=int(c.sum()),
               total_units=oracle.flow_cost(c, res["flow"]))
    path = os.path.join(HERE, "c5_whole_summary.npz")
    np.savez_compressed(path, **out)
    print("%-22s V=%d E=%d K=%d total=%.7f (%d KB)" % ("c5_whole_summary", out["n_nodes"], out["n_arcs"], res["K"],
                                                       res["total"], os.path.getsize(path) // 1024))


if __name__ == "__main__":
    main()
