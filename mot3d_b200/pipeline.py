"""TrackPipeline: the whole hot path on the device with preallocated buffers and no host synchronisation between
stages -- keys -> node costs -> frame sort -> single-pass arc build -> solve -> tracks (fused=False: arc count -> row
scan -> arc fill instead of the single-pass build). All graphs of a batch go
through each kernel in one launch (the solver drains a work queue of connected components). This is what the batch scheduler runs
per rank; bench.py times it stage by stage with CUDA events on the launching stream."""
import ctypes

import numpy as np
import torch

from . import _lib
from .batch import _ptr, _stream, _dets_struct, _views_mode, tracks_to_lists

STAGES = ["make_keys", "node_costs", "frame_sort", "edge_build", "solve", "extract"]
STAGES_UNFUSED = ["make_keys", "node_costs", "frame_sort", "edge_count", "row_scan", "edge_fill", "solve", "extract"]
# kernels launched by one run() (see the entry points in csrc/): keys 2, node costs 1, frame sort 1, single-pass
# build 1 (unfused: edge count 1, scan 3, edge fill + cost 2), solve 9 (components 3, first tree, single-track pass,
# job order, solver x2: giants + main, first-path rule), extract 1
LAUNCHES_PER_RUN = 2 + 1 + 1 + 1 + 9 + 1
LAUNCHES_PER_RUN_UNFUSED = 2 + 1 + 1 + 1 + 3 + 2 + 9 + 1


class TrackPipeline(object):
    def __init__(self, n_max, g_max, edge_capacity, device=None, dim=2, pruned=True, fused=True):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        dev = self.device
        self.n_max, self.g_max, self.edge_capacity = int(n_max), int(g_max), int(edge_capacity)
        L = _lib.lib()
        i32 = dict(dtype=torch.int32, device=dev)
        i64 = dict(dtype=torch.int64, device=dev)
        n, G, E = self.n_max, self.g_max, self.edge_capacity
        self.tkey = torch.empty(n, **i32)
        self.gbase = torch.empty(G, **i64)
        self.entry_cost = torch.empty(n, **i64)
        self.node_cost = torch.empty(n, **i64)
        self.exit_cost = torch.empty(n, **i64)
        self.row_count = torch.empty(n, **i32)
        self.pruned = pruned
        self.fused = bool(fused and pruned)
        self.stages = STAGES if self.fused else STAGES_UNFUSED
        self.launches_per_run = LAUNCHES_PER_RUN if self.fused else LAUNCHES_PER_RUN_UNFUSED
        self.build_ws_bytes = L.mot3d_edge_build_workspace_bytes(n)
        self.build_ws = torch.empty(self.build_ws_bytes, dtype=torch.uint8, device=dev)
        self.sorted_pos = torch.empty(dim * n, dtype=torch.float64, device=dev) if pruned else None
        self.sorted_perm = torch.empty(n, **i32) if pruned else None
        self.row_ptr = torch.empty(n + 1, **i32)
        self.scan_ws_bytes = L.mot3d_scan_workspace_bytes(n)
        self.scan_ws = torch.empty(self.scan_ws_bytes, dtype=torch.uint8, device=dev)
        self.col = torch.empty(E, **i32)
        self.cost = torch.empty(E, **i64)
        self.overflow = torch.zeros(1, **i32)
        self.next = torch.empty(n, **i32)
        self.prev = torch.empty(n, **i32)
        self.n_paths = torch.empty(G, **i32)
        self.total_cost = torch.empty(G, **i64)
        self.stats = torch.zeros(_lib.SOLVE_STATS * G, **i64)
        self.solve_ws_bytes = L.mot3d_solve_workspace_bytes(n, G, E)
        self.solve_ws = torch.empty(self.solve_ws_bytes, dtype=torch.uint8, device=dev)
        self.track_count = torch.empty(G, **i32)
        self.track_ptr = torch.empty(n + G + 1, **i32)
        self.track_det = torch.empty(n, **i32)
        self.weight = None   # float64 [E] arc weights, allocated when a cost with an appearance term comes through
        self.app_err = None

    def run(self, batch, spec, spec_c=None, events=None, before_extract=None, before_solve=None):
        """batch: DetectionBatch on self.device. events: optional list of len(self.stages)+1 torch.cuda.Event recorded
        around every stage on the current stream. before_extract: optional event the stream waits for before the
        track outputs are overwritten (a consumer of the previous run's outputs on another stream); before_solve: the
        same wait placed in front of the solver instead (its persistent blocks starve a concurrent NCCL kernel).
        Asynchronous: call check_overflow()/synchronize afterwards."""
        L = _lib.lib()
        dev = self.device
        n, G = batch.n, batch.n_graphs
        if n > self.n_max or G > self.g_max:
            raise ValueError("batch larger than the pipeline was sized for")
        cs = spec.to_c() if spec_c is None else spec_c
        views_mode = _views_mode(spec)
        cs_real = cs
        if views_mode:
            # 3-D detections with per-view appearance: the build leaves the distance factor in weight[] and
            # mot3d_edge_appearance_views finishes the arcs (as build_graphs does)
            cs = spec.to_c()
            cs.use_hist, cs.use_bbox, cs.defer_cost = 0, 0, 1
        need_w = bool(spec.use_hist or views_mode)
        if need_w and not self.fused:
            raise NotImplementedError("TrackPipeline(fused=False): the appearance passes need the single-pass build")
        if need_w and self.weight is None:
            self.weight = torch.empty(self.edge_capacity, dtype=torch.float64, device=dev)
            self.app_err = torch.zeros(1, dtype=torch.int32, device=dev)
        st = _stream(dev)
        rec = (lambda k: events[k].record(torch.cuda.current_stream(dev))) if events is not None else (lambda k: None)
        max_nodes = getattr(batch, "_max_nodes", None)
        if max_nodes is None:
            max_nodes = batch.max_graph_nodes()
        rec(0)
        _lib.check(L.mot3d_make_keys(_ptr(batch.frame), _ptr(batch.graph_ptr), G, spec.max_jump, _ptr(self.tkey),
                                     _ptr(self.gbase), _ptr(self.overflow), st))
        rec(1)
        d = _dets_struct(batch, self.tkey, spec, self.sorted_pos, self.sorted_perm)
        _lib.check(L.mot3d_node_costs(ctypes.byref(d), ctypes.byref(cs), _ptr(self.entry_cost), _ptr(self.node_cost),
                                      _ptr(self.exit_cost), st))
        rec(2)
        if self.pruned:
            _lib.check(L.mot3d_frame_sort(ctypes.byref(d), st))
        rec(3)
        if self.fused:
            _lib.check(L.mot3d_edge_build(ctypes.byref(d), ctypes.byref(cs), _lib.DIR_IN, 0, n, _ptr(self.row_ptr),
                                          self.edge_capacity, _ptr(self.col), _ptr(self.cost),
                                          _ptr(self.weight) if need_w else None,
                                          _ptr(self.overflow), None, None, _ptr(self.build_ws), self.build_ws_bytes,
                                          st))
            if views_mode:
                v = batch.views
                vs = _lib.Views()
                vs.n, vs.n_entries = n, int(v["view_id"].shape[0])
                vs.view_ptr, vs.view_id = v["view_ptr"].data_ptr(), v["view_id"].data_ptr()
                if spec.use_bbox:
                    vs.bbox = v["bbox"].data_ptr()
                if spec.use_hist:
                    vs.hist = v["hist"].data_ptr()
                    vs.hist_channels, vs.hist_bins = int(v["hist"].shape[1]), int(v["hist"].shape[2])
                _lib.check(L.mot3d_edge_appearance_views(ctypes.byref(d), ctypes.byref(vs), ctypes.byref(cs_real),
                                                         _lib.DIR_IN, _ptr(self.row_ptr), _ptr(self.col),
                                                         self.edge_capacity, _ptr(self.weight), _ptr(self.cost),
                                                         _ptr(self.app_err), st))
            elif spec.use_hist:
                _lib.check(L.mot3d_edge_appearance(ctypes.byref(d), ctypes.byref(cs), _lib.DIR_IN, _ptr(self.row_ptr),
                                                   _ptr(self.col), self.edge_capacity, _ptr(self.weight),
                                                   _ptr(self.cost), st))
            k = 4
        else:
            _lib.check(L.mot3d_edge_count(ctypes.byref(d), ctypes.byref(cs), _lib.DIR_IN, _ptr(self.row_count), st))
            rec(4)
            _lib.check(L.mot3d_exclusive_scan_i32(_ptr(self.row_count), n, _ptr(self.row_ptr), _ptr(self.scan_ws),
                                                  self.scan_ws_bytes, st))
            rec(5)
            _lib.check(L.mot3d_edge_fill(ctypes.byref(d), ctypes.byref(cs), _lib.DIR_IN, _ptr(self.row_ptr),
                                         self.edge_capacity, _ptr(self.col), _ptr(self.cost), None,
                                         _ptr(self.overflow), st))
            k = 6
        rec(k)
        if before_solve is not None:
            torch.cuda.current_stream(dev).wait_event(before_solve)
        _lib.check(L.mot3d_solve_batch(G, n, _ptr(batch.graph_ptr), max_nodes, _ptr(self.tkey), self.edge_capacity,
                                       _ptr(self.row_ptr),
                                       _ptr(self.col), _ptr(self.cost), _ptr(self.entry_cost), _ptr(self.node_cost),
                                       _ptr(self.exit_cost), _ptr(self.next), _ptr(self.prev), _ptr(self.n_paths),
                                       _ptr(self.total_cost), None, _ptr(self.stats), _ptr(self.solve_ws),
                                       self.solve_ws_bytes, st))
        rec(k + 1)
        if before_extract is not None:
            torch.cuda.current_stream(dev).wait_event(before_extract)
        _lib.check(L.mot3d_extract_tracks(G, _ptr(batch.graph_ptr), _ptr(self.next), _ptr(self.prev),
                                          _ptr(self.track_count), _ptr(self.track_ptr), _ptr(self.track_det), st))
        rec(k + 2)
        self._n, self._G = n, G

    def check_overflow(self):
        if self.app_err is not None and int(self.app_err.item()):
            raise ValueError("two linked 3-D detections share no 2-D view: weight_distance_detections_3d is nan for "
                             "them (reference mot3d/weight_functions.py:55-71)")
        flag = int(self.overflow.item())
        if flag & (_lib.ERR_UNSORTED | _lib.ERR_KEY_RANGE):
            from .batch import _raise_key_errors
            _raise_key_errors(flag)
        if flag != 0:
            raise _lib.Mot3dError(_lib.E_CAPACITY, "edge capacity %d too small (batch has %d arcs)" %
                                  (self.edge_capacity, int(self.row_ptr[self._n].item())))

    def n_edges(self):
        return int(self.row_ptr[self._n].item())

    def tracks(self, graph_ptr_host):
        return tracks_to_lists(graph_ptr_host, self.track_count.cpu().numpy(), self.track_ptr.cpu().numpy(),
                               self.track_det.cpu().numpy())
