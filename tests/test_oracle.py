"""CPU tests: pin the oracle (oracle/ssp_oracle.c + oracle/oracle.py) to the reference's own outputs
(tests/golden, produced by the unmodified reference Python + vendored muSSP) and, when oracle/_ref was built
from the vendored sources, to the real muSSP run live."""
import numpy as np
import pytest

from oracle import oracle
import golden_util as gu


@pytest.mark.parametrize("name", gu.GRAPH_CASES)
def test_ssp_restatement_matches_mussp(name):
    d = gu.load(name)
    r = oracle.ssp_solve(int(d["n_nodes"]), d["tail"], d["head"], d["cost"])
    assert r["K"] == int(d["K"])
    assert oracle.flow_cost(d["cost"], r["flow"]) == r["total"] == int(r["path_cost"].sum())
    if name in gu.MUSSP_BOGUS_COST:
        assert r["total"] == 5000000
    else:
        gap = gu.MUSSP_SUBOPTIMAL.get(name, 0)
        assert gu.ref_total_units(d) - r["total"] == gap  # exact optimum <= muSSP, equal where tolerances are silent
        if gap == 0:
            if name in gu.MUSSP_CLIPPED_PATH_COSTS:
                assert int(gu.ref_path_units(d).sum()) == r["total"]
            else:
                assert (r["path_cost"] == gu.ref_path_units(d)).all()
            assert (r["flow"] == d["flow"]).all()
    # successive path costs never decrease (convexity) and all but the first are negative
    assert (np.diff(r["path_cost"]) >= 0).all()
    assert (r["path_cost"][1:] < 0).all()


@pytest.mark.parametrize("name", gu.GRAPH_CASES)
def test_several_paths_per_search_equal_one_by_one(name):
    """The CUDA solver takes several paths from one set of labels (csrc/solve_sm.cuh "next path": sinks in ascending
    order while each lies in a first-level branch of the tree no earlier path of the batch went through). The rule,
    restated on the generic residual graph (ssp_oracle.c: oracle_ssp_solve_batched), must give exactly the paths of
    one-by-one successive shortest paths: same K, same path costs in the same order, same flow - with fewer searches."""
    d = gu.load(name)
    one = oracle.ssp_solve(int(d["n_nodes"]), d["tail"], d["head"], d["cost"])
    many = oracle.ssp_solve(int(d["n_nodes"]), d["tail"], d["head"], d["cost"], batched=True)
    assert many["K"] == one["K"] and many["total"] == one["total"]
    assert (many["path_cost"] == one["path_cost"]).all()
    assert (many["flow"] == one["flow"]).all()
    assert many["searches"] <= one["K"] + 1
    if name == "c5_window_f50_p100":
        assert many["searches"] < 15  # 100 paths


def test_several_paths_per_search_on_random_windows():
    """The same on 40 seeded crowded windows (people crossing, misses, clutter; costs in 1e-7 units): total, K and the
    sequence of path costs are those of one-by-one successive shortest paths, the flow costs the same."""
    n_batched = n_paths = 0
    for seed in range(40):
        rng = np.random.RandomState(seed)
        F, P = int(rng.randint(6, 25)), int(rng.randint(2, 12))
        pos = rng.rand(P, 2) * rng.choice([40.0, 120.0, 400.0])
        vel = rng.randn(P, 2) * 2.0
        frame, pts = [], []
        for f in range(F):
            pos = pos + vel + rng.randn(P, 2) * 0.5
            for q in range(P):
                if rng.rand() < 0.85:
                    frame.append(f)
                    pts.append(pos[q] + rng.randn(2))
            for _ in range(rng.poisson(0.5)):
                frame.append(f)
                pts.append(rng.rand(2) * 400.0)
        frame, pts = np.array(frame), np.array(pts).reshape(-1, 2)
        n = len(frame)
        spec = oracle.CostSpec(sigma_jump=1, sigma_distance=10, max_distance=30, use_color_histogram=False,
                               use_bbox=False, mul=1, bias=0, weight_source_sink=float(rng.choice([0.5, 1.0, 2.0])),
                               max_jump=int(rng.randint(1, 5)))
        conf = rng.uniform(0.3, 1.0, n)
        row_ptr, col, w = oracle.build_transitions(frame, pts, spec)
        tail, head, ww = oracle.graph_arcs(n, float(spec.weight_source_sink), oracle.node_weights(conf, spec), 0.0,
                                           row_ptr, col, w)
        cost = oracle.quantise(ww)
        one = oracle.ssp_solve(2 * n + 2, tail, head, cost)
        many = oracle.ssp_solve(2 * n + 2, tail, head, cost, batched=True)
        assert many["K"] == one["K"] and many["total"] == one["total"], seed
        assert (many["path_cost"] == one["path_cost"]).all(), seed
        assert oracle.flow_cost(cost, many["flow"]) == many["total"], seed
        n_batched += many["searches"]
        n_paths += one["K"]
    assert n_paths > 150 and n_batched < n_paths  # the batches do happen


@pytest.mark.parametrize("name", ["c1_simple_2d", "c4_pets_view1_w0", "c4_pets_view1_w1", "c4_synth_appearance",
                                  "c5_small_f30_p12", "wrapper_test_graph", "c2_tracklets_pass2"])
def test_exact_optimum_network_simplex(name):
    """Independent second oracle: min-cost circulation with a free T->S arc."""
    import networkx as nx
    d = gu.load(name)
    n = int(d["n_nodes"])
    g = nx.DiGraph()
    for t, h, c in zip(d["tail"].tolist(), d["head"].tolist(), d["cost"].tolist()):
        g.add_edge(t, h, weight=c, capacity=1)
    g.add_edge(n, 1, weight=0, capacity=10 ** 9)
    cost, flow = nx.network_simplex(g)
    r = oracle.ssp_solve(n, d["tail"], d["head"], d["cost"])
    assert cost == r["total"]
    assert flow[n][1] == r["K"]


@pytest.mark.parametrize("name", gu.DETECTION_CASES)
def test_build_restatement_matches_reference(name):
    d = gu.load(name)
    s = d["spec"]
    spec = oracle.CostSpec(**{k: v for k, v in s.items() if k != "kind"})
    row_ptr, col, w = oracle.build_transitions(d["frame"], d["pos"], spec, bbox=d.get("bbox"), hist=d.get("hist"),
                                               views=gu.views_of(d))
    n = len(d["frame"])
    tail, head, ww = oracle.graph_arcs(n, float(spec.weight_source_sink), oracle.node_weights(d["conf"], spec), 0.0,
                                       row_ptr, col, w)
    assert (tail == d["tail"]).all() and (head == d["head"]).all()  # same arcs, same order
    np.testing.assert_allclose(ww, d["w_float"], rtol=1e-12, atol=0)
    gu.assert_costs_match(oracle.quantise(ww), d["cost"], d["w_float"])
    assert (oracle.quantise(d["w_float"]) == d["cost"]).all()
    if name != "c5_window_f50_p100":  # text of 64k arcs: slow, sha checked on the others
        assert oracle.text_sha(oracle.arcs_to_text(2 * n + 2, tail, head, d["w_float"])) == str(d["sha"])


@pytest.mark.parametrize("name", gu.DETECTION_CASES)
def test_tracks_from_flow_matches_reference(name):
    d = gu.load(name)
    tr = oracle.tracks_from_flow(len(d["frame"]), d["tail"], d["head"], d["flow"])
    assert np.cumsum([0] + [len(t) for t in tr]).tolist() == d["track_ptr"].tolist()
    assert [i for t in tr for i in t] == d["track_det"].tolist()


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["wrapper_test_graph", "c1_simple_2d", "c3_soldiers_pass1", "c4_pets_view1_w1",
                                  "mussp_seq07"])
def test_live_mussp_reproduces_golden(name):
    d = gu.load(name)
    r = oracle.mussp_solve(int(d["n_nodes"]), d["tail"], d["head"], d["cost"].astype(np.float64) * 1e-7)
    assert r["K"] == int(d["K"])
    assert round(r["total"] * 1e7) == gu.ref_total_units(d)
    assert (r["flow"] == d["flow"]).all()


def test_c5_whole_summary():
    d = gu.load("c5_whole_summary")
    assert int(d["K"]) == 100 and int(d["n_arcs"]) == 1187969 and int(d["n_nodes"]) == 200002
    assert int(d["total_units"]) == gu.ref_total_units(d) == -1878765227978


@pytest.mark.parametrize("name", ["c2_tracklets_pass2", "c3_soldiers_pass2", "c4_synth_tracklets"])
def test_tracklet_build_restatement_matches_reference(name):
    d = gu.load(name)
    s = d["spec"]
    spec = oracle.TrackletSpec(**s)
    row_ptr, col, w = oracle.build_tracklet_transitions(d, spec)
    n = len(d["t_conf"])
    node_w = -(d["t_conf"] * spec.mul + spec.bias)
    tail, head, ww = oracle.graph_arcs(n, float(spec.weight_source_sink), node_w, 0.0, row_ptr, col, w)
    assert (tail == d["tail"]).all() and (head == d["head"]).all()
    np.testing.assert_allclose(ww, d["w_float"], rtol=1e-9, atol=0)
    gu.assert_costs_match(oracle.quantise(ww), d["cost"], d["w_float"])


def test_merge_close_points_restatement_matches_reference():
    """oracle.merge_close_points vs the reference's sklearn-DBSCAN version (mot3d/utils/filt.py:6-24) on the frames of
    the 3-D example and on a crowded synthetic set (tests/golden/c3_merge_close_points.npz)."""
    d = gu.load("c3_merge_close_points")
    rp, mp = d["raw_ptr"], d["merged_ptr"]
    for f in range(len(rp) - 1):
        got = oracle.merge_close_points(d["raw"][rp[f]:rp[f + 1]], eps=0.25)
        np.testing.assert_allclose(np.asarray(got).reshape(-1, 3), d["merged"][mp[f]:mp[f + 1]], rtol=1e-15, atol=0)
    groups = oracle.merge_close_points(d["crowd"], eps=0.3, return_indexes=True)
    gp = d["crowd_group_ptr"]
    assert [len(g) for g in groups] == np.diff(gp).tolist()
    assert [i for g in groups for i in g] == d["crowd_group_idx"].tolist()
    np.testing.assert_allclose(oracle.merge_close_points(d["crowd"], eps=0.3), d["crowd_merged"], rtol=1e-15, atol=0)
