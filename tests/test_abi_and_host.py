"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/mot3d_b200.h declares, argument
errors are reported through the ABI without touching CUDA, and the host-side logic (cost-spec recovery from
callables, quantisation, thresholds, packing order, scheduler + gather over gloo with world_size 2)."""
import ctypes
import functools
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    from mot3d_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "mot3d_b200.h")).read()
    declared = set(re.findall(r"\b(mot3d_[a-z0-9_]+)\s*\(", header))
    declared -= {"mot3d_last_error"} - {"mot3d_last_error"}
    assert len(declared) >= 14
    lib = ctypes.CDLL(built.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(built.SIGNATURES), declared ^ set(built.SIGNATURES)
    assert built.version().startswith("mot3d_b200")


def test_argument_errors_without_gpu(built):
    L = built.lib()
    assert L.mot3d_edge_count(None, None, 0, None, None) == built.E_INVALID
    assert b"NULL" in L.mot3d_last_error()
    d = built.Dets()
    d.n, d.dim = 4, 5
    s = built.CostSpecC()
    s.max_jump = 3
    assert L.mot3d_edge_count(ctypes.byref(d), ctypes.byref(s), 0, None, None) == built.E_INVALID
    assert b"dim" in L.mot3d_last_error()
    d.dim = 2
    s.max_jump = 100000
    assert L.mot3d_edge_fill(ctypes.byref(d), ctypes.byref(s), 0, None, 0, None, None, None, None, None) == built.E_INVALID
    assert b"max_jump" in L.mot3d_last_error()
    assert L.mot3d_solve_batch(-1, 0, None, 0, None, 0, None, None, None, None, None, None, None, None, None, None,
                               None, None, None, 0, None) == built.E_INVALID
    assert L.mot3d_scan_workspace_bytes(1 << 20) >= 8 * ((1 << 20) // 4096)
    assert L.mot3d_solve_workspace_bytes(1000, 3, 5000) >= 1000 * 40 + 5000 * 12
    # empty batches are fine and touch nothing
    d.n = 0
    s.max_jump = 3
    assert L.mot3d_edge_count(ctypes.byref(d), ctypes.byref(s), 0, None, None) == 0
    assert L.mot3d_extract_tracks(0, None, None, None, None, None, None, None) == 0


def test_spec_recovery_from_callables(built):
    import mot3d_b200 as mot3d
    import mot3d_b200.weight_functions as wf
    from mot3d_b200.mot3d import _recover_spec
    dets = [mot3d.Detection2D(0, (0, 0))]
    lam = lambda a, b: wf.weight_distance_detections_2d(a, b, sigma_jump=4, sigma_distance=10, max_distance=30,
                                                        use_color_histogram=False, use_bbox=True, sigma_box_size=0.1)
    conf = functools.partial(wf.weight_confidence_detections_2d, mul=0, bias=0.11)
    s = _recover_spec(dets, 1, 4, conf, lam)
    assert s.kind == "detections_2d" and s.distance["sigma_jump"] == 4 and s.distance["max_distance"] == 30
    assert s.use_bbox and not s.use_hist and s.confidence == dict(mul=0, bias=0.11)
    c = s.to_c()
    assert c.entry_cost == 10 ** 7 and c.max_jump == 4 and c.sigma_bbox_sq == 0.1 ** 2
    assert c.neg_exp_jump[0] == -1.0 and c.neg_exp_jump[3] == float(-np.exp(-3 * 4))
    # defaults of the bare functions (reference mot3d/mot3d.py:22-23)
    s = _recover_spec(dets, 1, 12, wf.weight_confidence_detections_2d, wf.weight_distance_detections_2d)
    assert s.distance["max_distance"] == 20 and s.use_hist and s.confidence == dict(mul=1, bias=0)
    # 3D detections, auto dispatch
    d3 = [mot3d.Detection3D(0, (0, 0, 0))]
    auto = lambda a, b: wf.weight_distance_detections_auto(a, b, config2d={}, config3d=dict(max_distance=0.4))
    s = _recover_spec(d3, 50, 8, lambda d: wf.weight_confidence_detections_auto(d), auto)
    assert s.kind == "detections_3d" and s.distance["max_distance"] == 0.4 and s.dim == 3
    with pytest.raises(TypeError):
        _recover_spec(dets, 1, 4, conf, lambda a, b: -1.0)
    with pytest.raises(TypeError):
        _recover_spec(dets, 1, 4, conf, lambda a, b: a.position[0] * b.nothing)
    with pytest.raises(TypeError):
        wf.weight_distance_detections_2d(dets[0], dets[0], bogus=1)


def test_quantisation_and_threshold(built):
    from mot3d_b200 import spec
    assert spec.quantise_scalar(0.1) == 1000000 and spec.quantise_scalar(-0.00000004) == 0
    assert spec.quantise_scalar(-0.00000005) == -1 or spec.quantise_scalar(-0.00000005) == 0  # binary value decides
    assert spec.quantise_scalar(50) == 500000000
    rng = np.random.RandomState(0)
    for m in [15, 30, 0.4, 20, 5, 1e-3, 12345.678] + rng.rand(50).tolist():
        s = spec.sq_dist_max(m)
        assert np.sqrt(np.float64(s)) <= m and np.sqrt(np.nextafter(np.float64(s), np.inf)) > m


def test_packing_order_is_the_reference_sort(built):
    import mot3d_b200 as mot3d
    frame = np.array([3, 1, 2, 1, 3, 0, 2, 1])
    pos = np.arange(16, dtype=float).reshape(8, 2)
    b, order = mot3d.DetectionBatch.from_arrays(np.array([0, 5, 8]), frame, pos)
    # graph 0 = items 0..4 sorted stably by frame, graph 1 = items 5..7
    assert order.tolist() == [1, 3, 2, 0, 4, 5, 7, 6]
    assert b.frame.tolist() == [1, 1, 2, 3, 3, 0, 1, 2] and b.pos.shape == (2, 8)
    assert b.pos[:, 0].tolist() == pos[1].tolist()


def test_types_api(built):
    import mot3d_b200 as mot3d
    a, b = mot3d.Detection2D(3, (1, 2)), mot3d.Detection2D(7, (2, 2), confidence=0.9)
    assert a.diff_index(b) == 4 and a.confidence == 0.5 and a.entry_points and a.exit_point and a.pre_node is None
    chain = [mot3d.Detection2D(i, (float(i), 0.0), bbox=(0, 0, 1 + i, 2)) for i in range(10)]
    t1, t2 = mot3d.DetectionTracklet2D(chain[:4], window=2), mot3d.DetectionTracklet2D(chain[6:])
    assert t1.diff_index(t2) == 6 - 3 and t1.head.indexes == [2, 3] and t1.tail.indexes == [0, 1]
    assert t2.tail.index == 6 and t2.head.position == (9.0, 0.0)
    with pytest.raises(ValueError):
        mot3d.DetectionTracklet3D(chain[:2])
    assert abs(mot3d.bbox_size_similarity((0, 0, 1, 1), (0, 0, 1, 1)) - 1) < 1e-12
    h = [np.ones(8)] * 3
    assert abs(mot3d.color_histogram_similarity(h, h) - 1) < 1e-12
    assert mot3d.linear_motion([0, 1], [(0, 0), (1, 0)], [2, 3], [(2, 0), (3, 0)]) < 1e-9


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under mot3d_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "mot3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("test oracle", ""), os.path.join(dirpath, f)


def test_no_gpu_means_loud_failure(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mot3d_b200 as mot3d
    with pytest.raises(RuntimeError):
        mot3d.build_graph([mot3d.Detection2D(0, (0, 0))], verbose=False,
                          weight_distance=lambda a, b: mot3d.weight_distance_detections_2d(a, b, use_color_histogram=False, use_bbox=False))


def test_lpt_sharding(built):
    import mot3d_b200 as mot3d
    sizes = [5000] * 20 + [300] * 16 + [26, 3385]
    for w in (1, 2, 4, 8):
        parts = mot3d.shard_graphs(sizes, w)
        assert sorted(np.concatenate(parts).tolist()) == list(range(len(sizes)))
        loads = [sum(sizes[g] ** 1.5 for g in p) for p in parts]
        assert max(loads) <= min(loads) + max(s ** 1.5 for s in sizes)
        assert [p.tolist() for p in parts] == [p.tolist() for p in mot3d.shard_graphs(sizes, w)]


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import mot3d_b200 as mot3d
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
sizes = [7, 3, 9, 4, 5]
parts = mot3d.shard_graphs(sizes, 2)
mine = parts[rank]
tracks = [[[g, g + 1, g + 2], [10 * g]] if g %% 2 == 0 else [] for g in mine]
costs = [-(g + 1) * 1000 for g in mine]
out = mot3d.gather_tracks(mine, tracks, costs, len(sizes))
if rank == 0:
    tr, co = out
    assert co == [-(g + 1) * 1000 for g in range(5)], co
    assert tr == [[[g, g + 1, g + 2], [10 * g]] if g %% 2 == 0 else [] for g in range(5)], tr
    print("GATHER_OK")
else:
    assert out is None
dist.destroy_process_group()
'''


def test_gather_tracks_gloo_world2(built, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % dict(root=ROOT, port=29500 + os.getpid() % 2000))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]


def test_track_payload_host_mirror_and_gloo_gather(built):
    """TrackGather over a gloo group of two processes with ragged shards (host mirror of the device payload)."""
    port = 31500 + os.getpid() % 2000
    worker = os.path.join(ROOT, "tests", "gather_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(port), "cpu"], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]


def test_trajectory_glue(built):
    import mot3d_b200 as mot3d
    traj = [mot3d.Detection2D(i, (i, 0)) for i in [3, 4, 5, 6, 9, 10, 11, 14, 15, 21]]
    pieces = mot3d.split_trajectory_modulo(traj, length=5)
    assert [[d.index for d in p] for p in pieces] == [[3, 4], [5, 6, 9], [10, 11, 14], [15, 21]]  # only a wrap of index % length cuts
    assert [[d.index for d in p] for p in mot3d.remove_short_trajectories(pieces, th_length=2)] == [[5, 6, 9], [10, 11, 14]]
    assert [d.index for d in mot3d.concat_tracklets(pieces)] == [d.index for d in traj]
    assert mot3d.split_trajectory_modulo([], length=5) == []


def test_interpolate_trajectory_and_scene_bookkeeping():
    """Host glue of the online loop (no GPU): hole filling like the reference's interpolate_trajectory
    (mot3d/utils/trajectory.py:14-69) and Scene's id / completion rules (projects/PETS2009S2L1/scene.py:8-40)."""
    import numpy as np
    import mot3d_b200 as mot3d
    from mot3d_b200.online import Scene
    a = mot3d.Detection2D(3, np.array([0.0, 10.0]), bbox=(0.0, 0.0, 2.0, 4.0))
    b = mot3d.Detection2D(7, np.array([8.0, 2.0]), bbox=(4.0, 4.0, 10.0, 12.0))
    c = mot3d.Detection2D(8, np.array([9.0, 2.0]), bbox=(5.0, 4.0, 11.0, 12.0))
    out = mot3d.interpolate_trajectory([a, b, c], attr_names=["position", "bbox"])
    assert [d.index for d in out] == [3, 4, 5, 6, 7, 8]
    assert out[0] is a and out[4] is b and out[5] is c
    np.testing.assert_allclose(out[2].position, [4.0, 6.0])
    np.testing.assert_allclose(out[1].bbox, [1.0, 1.0, 4.0, 6.0])
    assert type(out[1]) is mot3d.Detection2D
    scene = Scene(completed_after=5)
    scene.update(None, [a])
    scene.update(None, [b, c])
    scene.update(0, [a, b])          # existing id: replaced
    assert sorted(scene.active) == [0, 1] and scene.active[0] == [a, b]
    scene.update_completed(13)       # 7 + 5 < 13: trajectory 0 is done; 8 + 5 < 13 is false
    assert sorted(scene.completed) == [0] and sorted(scene.active) == [1]
    scene.update(None, [c])
    assert sorted(scene.active) == [1, 2]


def test_view_packing_follows_dict_order():
    """Host packing of the 2-D views of 3-D detections / tracklets (no GPU): entries in every object's own dict order
    (the order the reference's per-view loops iterate in, mot3d/weight_functions.py:55-71, 146-158), labels mapped to
    integers consistently, planar bboxes."""
    import types as pytypes
    import numpy as np
    import mot3d_b200 as mot3d
    from mot3d_b200.mot3d import _pack_views
    from mot3d_b200.tracklets import pack_tracklets

    def d2(view, k):
        return mot3d.Detection2D(0, (0.0, 0.0), view=view, color_histogram=[np.full(4, k + c, dtype=float) for c in range(3)],
                                 bbox=(k, k + 1.0, k + 2.0, k + 3.0))

    a = mot3d.Detection3D(0, (0.0, 0.0, 0.0), detections_2d={"camB": d2("camB", 1), "camA": d2("camA", 2)})
    b = mot3d.Detection3D(1, (0.0, 0.0, 0.0), detections_2d={"camA": d2("camA", 3)})
    c = mot3d.Detection3D(2, (0.0, 0.0, 0.0), detections_2d={})
    spec = mot3d.CostSpec(kind="detections_3d", distance=dict(use_color_histogram=True, use_bbox=True))
    v = _pack_views([a, b, c], spec)
    assert v["view_ptr"].tolist() == [0, 2, 3, 3]
    assert v["view_id"].tolist() == [0, 1, 1]                      # camB first seen -> 0, camA -> 1
    assert v["bbox"].shape == (4, 3) and v["bbox"][:, 0].tolist() == [1.0, 2.0, 3.0, 4.0]
    assert v["hist"].shape == (3, 3, 4) and v["hist"][1, 2, 0] == 4.0
    only_bbox = _pack_views([a, b], mot3d.CostSpec(kind="detections_3d", distance=dict(use_color_histogram=False,
                                                                                     use_bbox=True)))
    assert "hist" not in only_bbox and only_bbox["bbox"].shape == (4, 3)

    def end(idx, pos):
        e = mot3d.Detection3D(idx[-1], pos[-1])
        e.indexes, e.positions, e.access_point = idx, pos, False
        return e

    def trk(views):
        t = object.__new__(mot3d.DetectionTracklet3D)
        t._init_item(0.5, None)
        t.head = end([3, 4], [np.zeros(3), np.ones(3)])
        t.tail = end([0, 1], [np.zeros(3), np.ones(3)])
        t.tracklets_2d = {name: pytypes.SimpleNamespace(head=pytypes.SimpleNamespace(color_histogram=np.full((3, 4), h)),
                                                        tail=pytypes.SimpleNamespace(color_histogram=np.full((3, 4), -h)))
                          for name, h in views}
        return t

    tspec = mot3d.TrackletCostSpec(kind="tracklets_3d", distance=dict(use_color_histogram=True))
    packed = pack_tracklets([trk([("x", 1.0), ("y", 2.0)]), trk([("y", 3.0)])], tspec)
    assert packed["view_ptr"].tolist() == [0, 2, 3] and packed["view_id"].tolist() == [0, 1, 1]
    assert packed["view_head_hist"][2, 0, 0] == 3.0 and packed["view_tail_hist"][1, 0, 0] == -2.0
    assert packed["head_hist"] is None
    import pytest
    with pytest.raises(ValueError):
        t = trk([("x", 1.0)])
        t.tracklets_2d["x"].head.color_histogram = None
        pack_tracklets([t], tspec)


def test_dimacs_text_round_trip_on_files(built, tmp_path):
    """N4 at file level: save_graph (reference mot3d/mot3d.py:255-265) writes the lines graph_to_text made; reading the
    file back and parsing it (dimacs.parse_text, the text side of wrapper.cpp:39-88) gives the golden arcs with their
    exact 1e-7 costs, for a detection graph, a tracklet graph and muSSP's own sample."""
    import golden_util as gu
    from oracle import oracle
    import mot3d_b200 as mot3d
    from mot3d_b200 import dimacs
    for name in ("c1_simple_2d", "c2_tracklets_pass2", "a9_clipped_arcs", "wrapper_test_graph"):
        d = gu.load(name)
        w = d["w_float"] if "w_float" in d else d["cost"].astype(np.float64) * 1e-7
        lines = oracle.arcs_to_text(int(d["n_nodes"]), d["tail"], d["head"], w)
        path = str(tmp_path / (name + ".txt"))
        mot3d.save_graph(lines, path)
        back = open(path).read().splitlines()
        assert back == lines
        if "sha" in d and "w_float" in d:
            assert oracle.text_sha(back) == str(d["sha"])
        n_nodes, tail, head, cost = dimacs.parse_text(back)
        assert n_nodes == int(d["n_nodes"])
        assert (tail == d["tail"]).all() and (head == d["head"]).all() and (cost == d["cost"]).all()
    # signs, missing fraction digits, comments
    n, t, h, c = dimacs.parse_text(["p min 4 3", "c comment", "a 1 2 -0.5", "a 2 3 1", "a 3 4 +0.0000001"])
    assert n == 4 and c.tolist() == [-5000000, 10000000, 1]


def test_scene_save_load_round_trip(tmp_path):
    """Scene.save / Scene.load (reference projects/PETS2009S2L1/scene.py:41-52): the pickle carries active, completed
    and the tracklets under the reference's keys; the counters are rebuilt from the lengths."""
    import pickle
    import mot3d_b200 as mot3d
    from mot3d_b200.online import Scene
    sc = Scene(completed_after=5)
    for k in range(4):
        sc.update(None, [mot3d.Detection2D(f, (float(k), float(f)), confidence=0.5 + 0.1 * k) for f in range(k, k + 3)])
    sc.store_tracklet([mot3d.Detection2D(0, (1.0, 2.0))])
    sc.update_completed(time_index=8)      # tracks ending before frame 3 are completed
    assert sorted(sc.completed) == [0] and sorted(sc.active) == [1, 2, 3]
    path = str(tmp_path / "scene.pickle")
    sc.save(path)
    raw = pickle.load(open(path, "rb"))
    assert sorted(raw) == ["active", "completed", "trackelts"]
    other = Scene()
    other.load(path)
    assert sorted(other.active) == [1, 2, 3] and sorted(other.completed) == [0]
    assert other.id_last == 4 and other.id_last_tracklet == 1
    assert [d.index for d in other.active[2]] == [2, 3, 4] and other.active[3][0].confidence == 0.8
    assert other.tracklets[0][0].position == (1.0, 2.0)


def test_costs_on_real_items_and_input_checks(built):
    """weight_confidence_* on a real detection / tracklet is the reference's value (mot3d/weight_functions.py:37-38);
    packing refuses what the kernels cannot represent instead of truncating it."""
    import mot3d_b200 as mot3d
    import mot3d_b200.weight_functions as wf
    from mot3d_b200 import mot3d as m3
    from mot3d_b200.spec import CostSpec, quantise_scalar
    det = mot3d.Detection2D(3, (1.0, 2.0), confidence=0.83)
    assert wf.weight_confidence_detections(det, mul=2.0, bias=0.1) == -(0.83 * 2.0 + 0.1)
    assert wf.weight_confidence_detections_2d(det) == -0.83
    trk = mot3d.DetectionTracklet2D([det], confidence=0.4)
    assert wf.weight_confidence_tracklets_2d(trk, mul=1.5, bias=-0.2) == -(0.4 * 1.5 - 0.2)
    assert wf.weight_confidence_detections_auto(det, config2d=dict(mul=3)) == -(0.83 * 3)
    cs = CostSpec(kind="detections_2d")
    with pytest.raises(ValueError):
        m3._pack([mot3d.Detection3D(0, (1.0, 2.0, 3.0))], cs)          # z would be dropped
    with pytest.raises(ValueError):
        m3._pack([mot3d.Detection2D(0.5, (1.0, 2.0))], cs)             # fractional frame index
    with pytest.raises(ValueError):
        quantise_scalar(float("inf"))
    with pytest.raises(ValueError):
        quantise_scalar(float("nan"))
    with pytest.raises(NotImplementedError):
        wf.weight_distance_tracklets_auto(trk, mot3d.DetectionTracklet3D([mot3d.Detection3D(5, (0.0, 0.0, 0.0))]))


def test_options_and_release_without_gpu(built):
    """mot3d_set_option: known names take a value or NaN (= default), unknown names are an error with a message;
    mot3d_host_release on a thread that never ran the host call does nothing. No CUDA call involved."""
    L = built.lib()
    for name in ("host_chunks", "host_pieces", "host_split", "host_edge_weight", "host_trace", "host_no_pack",
                 "edge_cap", "edge_acap", "solve_no_fast", "big_delta"):
        assert L.mot3d_set_option(name.encode(), 1.0) == 0
        assert L.mot3d_set_option(name.encode(), float("nan")) == 0
    assert L.mot3d_set_option(b"no_such_option", 1.0) == built.E_INVALID
    assert b"no_such_option" in L.mot3d_last_error()
    assert L.mot3d_set_option(None, 1.0) == built.E_INVALID
    with pytest.raises(built.Mot3dError):
        built.set_option("nope", 3)
    assert L.mot3d_host_release() == 0
