"""Solver statistics on C5-whole-shaped graphs (one graph of F frames x 100 detections): solve time, rounds, pulls.
    python tools/c5whole_stats.py [frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import mot3d_b200 as mot3d
from window_latency import _gen
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
frame, pos = _gen(F, 100, 0)
spec = mot3d.CostSpec(kind="detections_2d", distance=dict(sigma_jump=1, sigma_distance=10, max_distance=30, use_color_histogram=False, use_bbox=False), confidence=dict(mul=1, bias=0), weight_source_sink=1, max_jump=8)
b, _ = mot3d.DetectionBatch.from_arrays(np.array([0, len(frame)]), frame, pos, np.full(len(frame), 0.9))
db = b.to("cuda")
gb = mot3d.build_graphs(db, spec)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    sol = mot3d.solve_graphs(gb)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
st = sol.stats.cpu().numpy()[:25]
names = ["asc_sweeps", "desc_sweeps", "node_pulls", "arc_pulls", "cyc_argmin", "cyc_list", "cyc_stage", "cyc_walk", "cyc_fix", "cyc_cert", "n_cert", "branches", "records_listed", "cyc_asc", "cyc_desc", "asc_steps", "desc_steps", "busy", "g_cyc_asc", "g_cyc_desc", "g_asc_steps", "g_desc_steps", "g_n", "g_cycles", "fast"]
print("F", F, "solve s", dt, "K", int(sol.n_paths[0]))
print({k: int(v) for k, v in zip(names, st)})
