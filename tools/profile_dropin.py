"""cProfile of the drop-in call (build_graph + solve_graph on Detection objects) on a small fixture: where the host
time of a small window goes.   python tools/profile_dropin.py [fixture] [calls]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu  # noqa: E402
import mot3d_b200 as mot3d  # noqa: E402
import mot3d_b200.weight_functions as wf  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c1_simple_2d"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 200
d = gu.load(name)
s = d["spec"]
n = len(d["frame"])
kw = {k: s[k] for k in ("sigma_jump", "sigma_distance", "sigma_color_histogram", "sigma_box_size", "max_distance",
                        "use_color_histogram", "use_bbox")}
hist, bbox = d.get("hist"), d.get("bbox")
dets = [mot3d.Detection2D(int(d["frame"][i]), d["pos"][i], confidence=float(d["conf"][i]),
                          color_histogram=None if hist is None else [h for h in hist[i]],
                          bbox=None if bbox is None else tuple(bbox[i])) for i in range(n)]
wd = lambda a, b: wf.weight_distance_detections_2d(a, b, **kw)  # noqa: E731
wc = lambda x: wf.weight_confidence_detections_2d(x, mul=s["mul"], bias=s["bias"])  # noqa: E731


def once():
    g = mot3d.build_graph(dets, weight_source_sink=s["weight_source_sink"], max_jump=s["max_jump"], verbose=False,
                          weight_confidence=wc, weight_distance=wd)
    return mot3d.solve_graph(g, verbose=False)


for _ in range(20):
    once()
t0 = time.perf_counter()
for _ in range(calls):
    once()
print("%s: %.3f ms per call" % (name, (time.perf_counter() - t0) / calls * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(calls):
    once()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
