"""Latency of ONE graph at a time (metric ii: build+solve ms per window), on the reference's own
configurations, beside the reference's CPU path on one host core. bench.py prints the result as its `latency` section;
stand-alone:

    python tools/window_latency.py [repeats]

Cases (fixtures made by the reference itself, tests/golden/, unless noted):
  C1 simple_2d, C3 soldiers_3d pass 1 with max_jump 8 (the example's value) and 4 (BASELINE.json's), the two real PETS
  windows (C4), one C5 window, C5-whole (synthetic 100 000 detections / 1 000 frames as ONE graph, generated), a long
  thin synthetic graph (F=1000, P=10, generated), the online loop of config 4 per batch (Tracker2D.update) and muSSP's
  own sample graph input_MOT_seq07 (solve only).
Per case:
  gpu_ms.objects   p50 / p99 of mot3d_b200.build_graph + solve_graph on Detection objects: the drop-in call, host packing,
                   H2D, kernels, D2H and unpacking included (small cases only)
  gpu_ms.arrays    p50 / p99 of DetectionBatch.from_arrays -> build_graphs -> solve_graphs -> extract_tracks -> host
                   lists: host arrays in, tracks out
  gpu_ms.host_call p50 / p99 of the C-ABI host-buffer call (HostTracker.load + mot3d_track_batch_host + tracks()): one
                   call per window, pinned buffers held by the caller (2-D cases without histograms)
  cpu_ms           the same graph through the numpy port of build_graph (oracle/oracle.py) + the reference's muSSP
                   (oracle/_ref, arrays in) on one core; `cpu_text_ms` = muSSP through the reference's text interface
                   (graph_to_text lines -> wrapper.cpp's parse -> solve: what mot3d itself pays per solve_graph call)
  totals agree     the GPU total cost equals muSSP's on the same arcs (checked on every case; not timed)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DIST_KEYS = ("sigma_jump", "sigma_distance", "sigma_color_histogram", "sigma_box_size", "max_distance",
             "use_color_histogram", "use_bbox")


def _gen(F, P, seed, extent=1000.0):
    """SURVEY.md 8(d) gen(F, P, seed) - the generator bench.py uses."""
    rng = np.random.RandomState(seed)
    pos = rng.rand(P, 2) * extent
    vel = rng.randn(P, 2) * 1.0
    out = np.empty((F, P, 2))
    for f in range(F):
        pos = pos + vel + rng.randn(P, 2) * 0.2
        out[f] = pos
    return np.repeat(np.arange(F), P), out.reshape(-1, 2)


def _pct(t):
    return {"p50": float(np.percentile(t, 50)), "p99": float(np.percentile(t, 99)), "n": len(t)}


def _time(fn, reps, budget_s):
    fn()
    t, t_start = [], time.perf_counter()
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append((time.perf_counter() - t0) * 1e3)
        if time.perf_counter() - t_start > budget_s:
            break
    return t


def measure(reps=50, budget_s=6.0, cpu=True, log=None):
    import torch
    import golden_util as gu
    import mot3d_b200 as mot3d
    import mot3d_b200.weight_functions as wf
    from mot3d_b200 import dimacs
    from oracle import oracle
    oracle.build()
    have_ref = oracle.have_ref()
    dev = torch.device("cuda", torch.cuda.current_device())
    out = []

    def detection_case(label, frame, pos, conf, s, hist=None, bbox=None, objects=True):
        n = len(frame)
        kw = {k: s[k] for k in DIST_KEYS}
        kind = "detections_2d" if s["kind"] == "2d" else "detections_3d"
        spec = mot3d.CostSpec(kind=kind, distance=kw, confidence=dict(mul=s["mul"], bias=s["bias"]),
                              weight_source_sink=s["weight_source_sink"], max_jump=s["max_jump"])
        gp = np.array([0, n])
        res = {}

        def arrays_once():
            b, _ = mot3d.DetectionBatch.from_arrays(gp, frame, pos, conf, bbox=bbox, hist=hist)
            gb = mot3d.build_graphs(b.to(dev), spec)
            sol = mot3d.solve_graphs(gb, want_path_costs=False)
            mot3d.extract_tracks(gb, sol)
            tr = mot3d.tracks_to_lists(gp, sol.track_count.cpu().numpy(), sol.track_ptr.cpu().numpy(),
                                       sol.track_det.cpu().numpy())
            res["total"], res["K"], res["arcs"] = int(sol.total_cost[0].item()), len(tr[0]), gb.n_edges
            return tr

        rec = {"case": label, "detections": n, "max_jump": s["max_jump"]}
        gpu = {"arrays": _pct(_time(arrays_once, reps, budget_s))}
        if hist is None and s["kind"] == "2d" and n <= 20000:
            # the C-ABI host-buffer call (mot3d_track_batch_host through HostTracker): load into pinned buffers, one call,
            # tracks out; sized once (arc capacity from the array path above), as a service would hold it
            ht = mot3d.HostTracker(n_max=n, g_max=1, edge_capacity=int(res["arcs"]) + 64, dim=2, with_bbox=bbox is not None)
            hb, _ = mot3d.DetectionBatch.from_arrays(gp, frame, pos, conf, bbox=bbox)
            spec_c = spec.to_c()

            def host_once():
                ht.load(hb)
                ht.run(spec_c)
                return ht.tracks()
            gpu["host_call"] = _pct(_time(host_once, reps, budget_s))
            res["host_total"] = int(ht.h_total[0].item())
        if objects:
            if s["kind"] == "3d":
                dets = [mot3d.Detection3D(int(frame[i]), pos[i], confidence=float(conf[i])) for i in range(n)]
                wd = lambda a, b: wf.weight_distance_detections_3d(a, b, **kw)  # noqa: E731
                wc = lambda x: wf.weight_confidence_detections_3d(x, mul=s["mul"], bias=s["bias"])  # noqa: E731
            else:
                dets = [mot3d.Detection2D(int(frame[i]), pos[i], confidence=float(conf[i]),
                                          color_histogram=None if hist is None else [h for h in hist[i]],
                                          bbox=None if bbox is None else tuple(bbox[i])) for i in range(n)]
                wd = lambda a, b: wf.weight_distance_detections_2d(a, b, **kw)  # noqa: E731
                wc = lambda x: wf.weight_confidence_detections_2d(x, mul=s["mul"], bias=s["bias"])  # noqa: E731

            def objects_once():
                g = mot3d.build_graph(dets, weight_source_sink=s["weight_source_sink"], max_jump=s["max_jump"],
                                      verbose=False, weight_confidence=wc, weight_distance=wd)
                return mot3d.solve_graph(g, verbose=False)
            gpu["objects"] = _pct(_time(objects_once, reps, budget_s))
        rec.update(arcs=int(res["arcs"]), tracks=int(res["K"]), gpu_ms=gpu)
        if cpu:
            ospec = oracle.CostSpec(**{k: v for k, v in s.items() if k != "kind"})
            t_arr, t_txt, ref_total = [], [], None
            for _ in range(3 if n < 20000 else 1):
                t0 = time.perf_counter()
                row_ptr, col, w = oracle.build_transitions(frame, pos, ospec, bbox=bbox, hist=hist)
                tail, head, ww = oracle.graph_arcs(n, float(ospec.weight_source_sink), oracle.node_weights(conf, ospec),
                                                   0.0, row_ptr, col, w)
                t_build = time.perf_counter() - t0
                t0 = time.perf_counter()
                if have_ref:
                    r = oracle.mussp_solve(2 * n + 2, tail, head, np.round(ww, 7))
                else:
                    r = oracle.ssp_solve(2 * n + 2, tail, head, oracle.quantise(ww))
                oracle.tracks_from_flow(n, tail, head, r["flow"])
                t_arr.append((t_build + time.perf_counter() - t0) * 1e3)
                if have_ref and n <= 20000:
                    lines = oracle.arcs_to_text(2 * n + 2, tail, head, ww)  # graph_to_text, not timed
                    t0 = time.perf_counter()
                    oracle.mussp_solve_text(lines, len(tail))
                    t_txt.append((t_build + time.perf_counter() - t0) * 1e3)
                ref_total = oracle.flow_cost(oracle.quantise(ww), r["flow"])
            rec["cpu_ms"] = {"p50": float(np.median(t_arr)), "cores": 1,
                             "what": "numpy port of build_graph + %s, arrays in" % ("muSSP (oracle/_ref)" if have_ref
                                                                                  else "C port of the solve")}
            if t_txt:
                rec["cpu_text_ms"] = {"p50": float(np.median(t_txt)), "cores": 1,
                                      "what": "numpy port of build_graph + muSSP through the text interface "
                                              "(wrapper.cpp:39-88 parse + solve)"}
            # muSSP's float tolerances may leave it a few units above the exact optimum
            rec["total_cost_units"] = {"gpu": int(res["total"]), "mussp_same_arcs": int(ref_total)}
            rec["totals_agree"] = bool(0 <= int(ref_total) - int(res["total"]) <= 2 and
                                       res.get("host_total", res["total"]) == res["total"])
        if log:
            log(rec)
        out.append(rec)

    for name, objects in (("c1_simple_2d", True), ("c4_pets_view1_w0", True), ("c4_pets_view1_w1", True),
                          ("c3_soldiers_pass1", True), ("c5_window_f50_p100", False)):
        d = gu.load(name)
        detection_case(name, d["frame"], d["pos"], d["conf"], d["spec"], hist=d.get("hist"), bbox=d.get("bbox"),
                       objects=objects)
    d = gu.load("c3_soldiers_pass1")
    s4 = dict(d["spec"], max_jump=4)
    detection_case("c3_soldiers_pass1_max_jump4", d["frame"], d["pos"], d["conf"], s4, objects=False)
    s5 = dict(gu.load("c5_window_f50_p100")["spec"])
    frame, pos = _gen(1000, 10, 7)
    detection_case("long_thin_f1000_p10", frame, pos, np.full(len(frame), 0.9), s5, objects=False)
    frame, pos = _gen(1000, 100, 0)
    detection_case("c5_whole_f1000_p100", frame, pos, np.full(len(frame), 0.9), s5, objects=False)

    # the online window loop (BASELINE config 4): Tracker2D.update per batch of 50 frames on C4-synth, the scene carried
    # from batch to batch (two graph builds + solves per batch and the host glue between them)
    import json as _json
    from mot3d_b200.online import Scene, Tracker2D
    d = gu.load("c4_online_loop")
    cfg = _json.loads(str(d["config"]))
    ptr, det = d["det_ptr"], d["det"]
    B = cfg["batch_size"]
    t_batches = []
    for rep in range(3):
        scene = Scene(cfg["completed_after"])
        tracker = Tracker2D(scene, None, cfg, image_size=(768, 576))
        for b in range((len(ptr) - 1) // B):
            dets = [[mot3d.Detection2D(f, det[r, :2].copy(), confidence=float(det[r, 2]), bbox=tuple(det[r, 3:7]))
                     for r in range(int(ptr[f]), int(ptr[f + 1]))] for f in range(b * B, (b + 1) * B)]
            t0 = time.perf_counter()
            tracker.update((b + 1) * B, dets)
            if rep:
                t_batches.append((time.perf_counter() - t0) * 1e3)
    rec = {"case": "c4_online_loop_per_batch", "detections": int(len(det) // ((len(ptr) - 1) // B)),
           "gpu_ms": {"update": _pct(t_batches)},
           "what": "mot3d_b200.online.Tracker2D.update: detection pass, tracklet pass against the active trajectories, "
                   "interpolation, scene update; 50 frames per batch",
           "totals_agree": bool(len(scene.completed) == 2 and len(scene.active) == 6)}
    # the same loop with the scene kept on the device (DeviceTracker2D: arrays in, no objects; N1)
    from mot3d_b200.online_device import DeviceTracker2D
    t_dev = []
    for rep in range(3):
        trk = DeviceTracker2D(cfg, completed_after=cfg["completed_after"])
        for b in range((len(ptr) - 1) // B):
            lo, hi = int(ptr[b * B]), int(ptr[(b + 1) * B])
            fr = np.repeat(np.arange(b * B, (b + 1) * B), np.diff(ptr[b * B:(b + 1) * B + 1]))
            t0 = time.perf_counter()
            trk.update((b + 1) * B, fr, det[lo:hi, :2], det[lo:hi, 2], det[lo:hi, 3:7])
            torch.cuda.synchronize()
            if rep:
                t_dev.append((time.perf_counter() - t0) * 1e3)
    rec["gpu_ms"]["update_device_scene"] = _pct(t_dev)
    rec["totals_agree"] = bool(rec["totals_agree"] and len(trk.completed) == 2 and len(trk.ids) == 6)
    if log:
        log(rec)
    out.append(rec)

    # muSSP's own sample graph (solve only): arcs in, flow out
    d = gu.load("mussp_seq07")
    tail, head, cost = d["tail"], d["head"], d["cost"]
    n_nodes = int(d["n_nodes"])
    gb, perm = dimacs.graph_from_arcs(n_nodes, tail, head, cost, device=dev)  # format conversion on the host: not timed
    res = {}

    def seq07_once():
        sol = mot3d.solve_graphs(gb, want_path_costs=False)
        res["total"] = int(sol.total_cost[0].item())  # (synchronises)

    rec = {"case": "mussp_seq07_solve_only", "detections": len(perm), "arcs": int(len(tail)),
           "gpu_ms": {"solve_device_arrays": _pct(_time(seq07_once, reps, budget_s))}}
    if cpu and have_ref:
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            r = oracle.mussp_solve(n_nodes, tail, head, cost.astype(np.float64) * 1e-7)
            t.append((time.perf_counter() - t0) * 1e3)
        ref_total = int(oracle.flow_cost(cost, r["flow"]))
        rec["cpu_ms"] = {"p50": float(np.median(t)), "cores": 1, "what": "muSSP (oracle/_ref), arrays in, solve only"}
        rec["total_cost_units"] = {"gpu": res["total"], "mussp_same_arcs": ref_total}
        rec["totals_agree"] = bool(0 <= ref_total - res["total"] <= 2)
    if log:
        log(rec)
    out.append(rec)
    return out


if __name__ == "__main__":
    measure(reps=int(sys.argv[1]) if len(sys.argv) > 1 else 50, log=lambda r: print(json.dumps(r), flush=True))
