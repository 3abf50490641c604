"""Profiling driver (run under ncu on the GPU box): a few C5 windows through the device pipeline, twice."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mot3d_b200 as mot3d  # noqa: E402
from mot3d_b200.pipeline import TrackPipeline  # noqa: E402
from bench import make_workload, DIST_KW, MAX_JUMP  # noqa: E402

n_graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
graph_ptr, frame, pos, conf = make_workload((n_graphs + 19) // 20, seed0=0)
n = int(graph_ptr[n_graphs])
graph_ptr = graph_ptr[:n_graphs + 1]
spec = mot3d.CostSpec(kind="detections_2d", distance=DIST_KW, confidence=dict(mul=1, bias=0), weight_source_sink=1,
                      max_jump=MAX_JUMP)
hb, _ = mot3d.DetectionBatch.from_arrays(graph_ptr, frame[:n], pos[:n], conf[:n])
db = hb.to("cuda")
gb = mot3d.build_graphs(db, spec)
pipe = TrackPipeline(n, n_graphs, gb.n_edges + 16)
sc = spec.to_c()
for _ in range(2):
    pipe.run(db, spec, sc)
torch.cuda.synchronize()
print("ok", int(pipe.n_paths.cpu().numpy()[:n_graphs].sum()))
