"""Timeline of the software pipeline inside mot3d_track_batch_host (run on the GPU box):
    python tools/host_pipeline_trace.py [sequences] [chunks]
prints, per chunk, when its H2D copies, build, solve+tracks and D2H copies were done (CUDA events, ms from call)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mot3d_b200 as mot3d  # noqa: E402
from bench import make_workload, DIST_KW, MAX_JUMP  # noqa: E402

seqs = int(sys.argv[1]) if len(sys.argv) > 1 else 60
if len(sys.argv) > 2:
    mot3d._lib.set_option("host_chunks", int(sys.argv[2]))
graph_ptr, frame, pos, conf = make_workload(seqs, seed0=0)
spec = mot3d.CostSpec(kind="detections_2d", distance=DIST_KW, confidence=dict(mul=1, bias=0), weight_source_sink=1,
                      max_jump=MAX_JUMP)
hb, _ = mot3d.DetectionBatch.from_arrays(graph_ptr, frame, pos, conf)
n, G = hb.n, hb.n_graphs
ht = mot3d.HostTracker(n_max=n, g_max=G, edge_capacity=int(n * 8.5), dim=2)
ht.load(hb)
sc = spec.to_c()
for _ in range(3):
    ht.run(sc)
ts = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ht.run(sc)
    ts.append((time.perf_counter() - t0) * 1e3)
print("call ms:", " ".join("%.2f" % t for t in ts))
mot3d._lib.set_option("host_trace", 1)
ht.run(sc)
