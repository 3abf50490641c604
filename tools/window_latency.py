"""Latency of ONE window through the drop-in surface (metric ii: build+solve ms per window), on the
fixtures of the reference's own configs: C1 simple_2d, C3 soldiers_3d pass 1, the two real PETS windows (C4) and one
C5 window. Prints a JSON line per case: p50 / p99 over repeated calls of build_graph + solve_graph on Detection
objects (host packing, H2D, kernels, D2H, unpacking all included) and, beside it, the same window through the numpy
port of build_graph + the vendored muSSP on one host core (oracle/, the CPU baseline of bench.py).
    python tools/window_latency.py [repeats]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu  # noqa: E402
import mot3d_b200 as mot3d  # noqa: E402
import mot3d_b200.weight_functions as wf  # noqa: E402
from oracle import oracle  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
oracle.build()
for name in ["c1_simple_2d", "c4_pets_view1_w0", "c4_pets_view1_w1", "c3_soldiers_pass1", "c5_window_f50_p100"]:
    d = gu.load(name)
    s = d["spec"]
    n = len(d["frame"])
    three_d = s["kind"] == "3d"
    kw = {k: s[k] for k in ("sigma_jump", "sigma_distance", "sigma_color_histogram", "sigma_box_size", "max_distance",
                            "use_color_histogram", "use_bbox")}
    if three_d:
        dets = [mot3d.Detection3D(int(d["frame"][i]), d["pos"][i], confidence=float(d["conf"][i])) for i in range(n)]
        wd = lambda a, b: wf.weight_distance_detections_3d(a, b, **kw)
        wc = lambda x: wf.weight_confidence_detections_3d(x, mul=s["mul"], bias=s["bias"])
    else:
        hist, bbox = d.get("hist"), d.get("bbox")
        dets = [mot3d.Detection2D(int(d["frame"][i]), d["pos"][i], confidence=float(d["conf"][i]),
                                  color_histogram=None if hist is None else [h for h in hist[i]],
                                  bbox=None if bbox is None else tuple(bbox[i])) for i in range(n)]
        wd = lambda a, b: wf.weight_distance_detections_2d(a, b, **kw)
        wc = lambda x: wf.weight_confidence_detections_2d(x, mul=s["mul"], bias=s["bias"])

    def once():
        g = mot3d.build_graph(dets, weight_source_sink=s["weight_source_sink"], max_jump=s["max_jump"], verbose=False,
                              weight_confidence=wc, weight_distance=wd)
        return mot3d.solve_graph(g, verbose=False)

    for _ in range(3):
        tr = once()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        once()
        t.append((time.perf_counter() - t0) * 1e3)
    ospec = oracle.CostSpec(**{k: v for k, v in s.items() if k != "kind"})
    c = []
    for _ in range(max(3, reps // 10)):
        t0 = time.perf_counter()
        row_ptr, col, w = oracle.build_transitions(d["frame"], d["pos"], ospec, bbox=d.get("bbox"), hist=d.get("hist"))
        tail, head, ww = oracle.graph_arcs(n, float(ospec.weight_source_sink), oracle.node_weights(d["conf"], ospec), 0.0,
                                           row_ptr, col, w)
        if oracle.have_ref():
            r = oracle.mussp_solve(2 * n + 2, tail, head, np.round(ww, 7))
        else:
            r = oracle.ssp_solve(2 * n + 2, tail, head, oracle.quantise(ww))
        oracle.tracks_from_flow(n, tail, head, r["flow"])
        c.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"window": name, "detections": n, "arcs": int(len(d["tail"])), "tracks": len(tr),
                      "gpu_build_plus_solve_ms": {"p50": float(np.percentile(t, 50)), "p99": float(np.percentile(t, 99)),
                                                  "api": "mot3d_b200.build_graph + solve_graph on Detection objects"},
                      "cpu_numpy_port_plus_mussp_ms": {"p50": float(np.percentile(c, 50)), "cores": 1}}))
