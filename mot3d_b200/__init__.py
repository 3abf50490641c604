"""mot3d_b200: the mot3d tracking hot path (graph build + min-cost-flow solve + trajectories) on B200.

Drop-in for `import mot3d` on that path: build_graph, solve_graph, graph_to_text, save_graph, the Detection* types,
weight_functions and distance_functions keep the reference's names (reference mot3d/__init__.py:1-7). The work runs
in hand-written sm_100a CUDA kernels behind a C ABI (include/mot3d_b200.h); importing this package does not need a
GPU, calling build_graph / solve_graph does, and there is no CPU fallback.
"""
from .types import *            # noqa: F401,F403
from .weight_functions import *  # noqa: F401,F403
from .distance_functions import *  # noqa: F401,F403
from .mot3d import *            # noqa: F401,F403
from .utils.trajectory import *  # noqa: F401,F403
from . import utils  # noqa: F401
from .mot3d import TrackingGraph, TrackletGraph  # noqa: F401
from . import weight_functions, distance_functions, types  # noqa: F401
from .spec import CostSpec, TrackletCostSpec  # noqa: F401
from .batch import (DetectionBatch, GraphBatch, HostTracker, build_graphs, solve_graphs, extract_tracks,  # noqa: F401
                    track_batch, tracks_to_lists)
from .scheduler import shard_graphs, gather_tracks, TrackGather, pack_tracks_host, unpack_tracks_host  # noqa: F401
from . import online  # noqa: F401
from .twopass import link_tracks  # noqa: F401
from .online_device import DeviceTracker2D  # noqa: F401
from ._lib import Mot3dError, version  # noqa: F401
