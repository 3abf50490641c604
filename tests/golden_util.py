"""Helpers shared by the parity tests: load tests/golden/*.npz (made by tests/golden/make_golden.py from the
reference itself) and compare quantised costs with the one tolerance the domain needs."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCALE = 10 ** 7

# Fixtures with detections (frame/pos/...) -> build + solve parity; the others are solver-only graphs.
DETECTION_CASES = ["c1_simple_2d", "c3_soldiers_pass1", "c3_synth_multiview", "c4_pets_view1_w0", "c4_pets_view1_w1",
                   "c4_synth_appearance", "c5_small_f30_p12", "c5_window_f50_p100", "quirk_single_detection",
                   "a9_clipped_arcs"]  # a9: muSSP's invalid_edge_rm clips 1 571 of its 1 805 transition arcs
TRACKLET_CASES = ["c2_tracklets_pass2", "c3_soldiers_pass2", "c3_synth_multiview_tracklets", "c4_synth_tracklets"]
GRAPH_CASES = DETECTION_CASES + TRACKLET_CASES + ["wrapper_test_graph", "mussp_seq07"]

# muSSP works in doubles with tolerances (Graph.cpp:175 '> 1e-7', :367 '< 1e-6'); on this window it leaves an
# arc of cost -1e-7 unused and ends 1 unit above the exact optimum (confirmed with networkx.network_simplex:
# -1052113361 vs muSSP -1052113360). The integer solvers (oracle and GPU) return the exact optimum.
MUSSP_SUBOPTIMAL = {"c4_pets_view1_w0": 1}
# First path is kept even when its cost is >= 0 (wrapper.cpp:205,220); muSSP then reports cost 0 for it because
# the sink's distance starts at its first in-arc weight (Graph.cpp:151). Integer solvers report the true +0.5.
MUSSP_BOGUS_COST = {"quirk_single_detection"}
# Graph::invalid_edge_rm (Graph.cpp:86-111) removes arcs that cost more than ending one track and starting another
# before muSSP solves: the optimum (flow, K, total) is untouched - no optimal flow uses such an arc - but the successive
# shortest paths muSSP reports on the way are those of the clipped graph, not of the graph it was given.
MUSSP_CLIPPED_PATH_COSTS = {"a9_clipped_arcs"}


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    if "spec" in d:
        d["spec"] = json.loads(str(d["spec"]))
    return d


def views_of(d):
    """The 2-D views of a fixture of 3-D detections (entries in each detection's dict order), or None:
    dict(view_ptr, view_id, bbox [E, 4], hist [E, C, B])."""
    if "view_ptr" not in d:
        return None
    return dict(view_ptr=d["view_ptr"], view_id=d["view_id"], bbox=d["view_bbox"], hist=d["view_hist"])


def all_cases():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz")))


def ref_total_units(d):
    return int(round(float(d["total"]) * SCALE))


def ref_path_units(d):
    return np.round(np.asarray(d["path_cost"], dtype=np.float64) * SCALE).astype(np.int64)


def assert_costs_match(got_units, ref_units, ref_w):
    """Quantised costs must be identical, except that a weight whose exact value lies within 1e-8 units of a
    rounding tie (x.5 units) may land on either neighbour: exp() differs by <= 1 ulp between numpy and CUDA, which
    is <= 2e-9 units on weights of magnitude <= 1 (1e7 units)."""
    got_units = np.asarray(got_units, dtype=np.int64)
    ref_units = np.asarray(ref_units, dtype=np.int64)
    assert got_units.shape == ref_units.shape
    bad = np.flatnonzero(got_units != ref_units)
    for e in bad.tolist():
        x = float(ref_w[e]) * SCALE
        frac = abs(x - np.floor(x) - 0.5)
        assert abs(int(got_units[e]) - int(ref_units[e])) == 1 and frac < 1e-8, (
            "arc %d: got %d ref %d (w=%r)" % (e, got_units[e], ref_units[e], ref_w[e]))
    return len(bad)
