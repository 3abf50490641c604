"""ctypes binding of libmot3d_b200.so (include/mot3d_b200.h). No fallback: if the CUDA library is missing or a call
fails, this raises. PyTorch only supplies device memory and the stream."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmot3d_b200.so")

MAX_JUMP = 256
INF_COST = 1 << 60
DIR_IN, DIR_OUT = 0, 1
SOLVE_BEGIN, SOLVE_FRONT, SOLVE_BACK = 1, 2, 4  # MOT3D_SOLVE_*: phases of mot3d_solve_batch_phase
SOLVE_STATS = 25  # include/mot3d_b200.h: MOT3D_SOLVE_STATS
FLAG_ENTRY, FLAG_EXIT = 1, 2
TRACK_START_BIT = 0x80000000
PEER_HANDLE_BYTES = 64

E_INVALID, E_CUDA, E_CAPACITY, E_UNSUPPORTED = -1, -2, -3, -4
ERR_ARC_CAPACITY, ERR_WINDOW, ERR_UNSORTED, ERR_KEY_RANGE = 1, 2, 4, 8  # MOT3D_ERR_*: bits of the build's flag word


class Mot3dError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "mot3d_b200 error %d: %s" % (code, msg))
        self.code = code


class Dets(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("dim", ctypes.c_int32), ("hist_channels", ctypes.c_int32),
                ("hist_bins", ctypes.c_int32), ("reserved", ctypes.c_int32), ("tkey", ctypes.c_void_p),
                ("pos", ctypes.c_void_p), ("conf", ctypes.c_void_p), ("flags", ctypes.c_void_p),
                ("bbox", ctypes.c_void_p), ("hist", ctypes.c_void_p), ("sorted_pos", ctypes.c_void_p),
                ("sorted_perm", ctypes.c_void_p)]


class CostSpecC(ctypes.Structure):
    _fields_ = [("max_jump", ctypes.c_int32), ("use_bbox", ctypes.c_int32), ("use_hist", ctypes.c_int32),
                ("defer_cost", ctypes.c_int32), ("sq_dist_max", ctypes.c_double),
                ("sigma_distance_sq", ctypes.c_double), ("sigma_hist_sq", ctypes.c_double),
                ("sigma_bbox_sq", ctypes.c_double), ("conf_mul", ctypes.c_double), ("conf_bias", ctypes.c_double),
                ("entry_cost", ctypes.c_int64), ("exit_cost", ctypes.c_int64),
                ("neg_exp_jump", ctypes.c_double * MAX_JUMP)]


class Views(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("n_entries", ctypes.c_int64), ("hist_channels", ctypes.c_int32),
                ("hist_bins", ctypes.c_int32), ("view_ptr", ctypes.c_void_p), ("view_id", ctypes.c_void_p),
                ("bbox", ctypes.c_void_p), ("hist", ctypes.c_void_p)]


class Tracklets(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("dim", ctypes.c_int32), ("hist_channels", ctypes.c_int32),
                ("hist_bins", ctypes.c_int32), ("n_head_points", ctypes.c_int64), ("n_tail_points", ctypes.c_int64),
                ("head_ptr", ctypes.c_void_p), ("tail_ptr", ctypes.c_void_p), ("head_idx", ctypes.c_void_p),
                ("tail_idx", ctypes.c_void_p), ("head_pos", ctypes.c_void_p), ("tail_pos", ctypes.c_void_p),
                ("tail_access_point", ctypes.c_void_p), ("head_has_hist", ctypes.c_void_p),
                ("tail_has_hist", ctypes.c_void_p), ("head_hist", ctypes.c_void_p), ("tail_hist", ctypes.c_void_p),
                ("head_fit_ws", ctypes.c_void_p), ("tail_fit_ws", ctypes.c_void_p),
                ("view_ptr", ctypes.c_void_p), ("view_id", ctypes.c_void_p), ("view_head_hist", ctypes.c_void_p),
                ("view_tail_hist", ctypes.c_void_p)]


class TrackletSpecC(ctypes.Structure):
    _fields_ = [("max_jump", ctypes.c_int32), ("use_hist", ctypes.c_int32), ("check_access_point", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("sq_dist_max", ctypes.c_double), ("sigma_motion_sq", ctypes.c_double),
                ("sigma_hist_sq", ctypes.c_double), ("alpha", ctypes.c_double), ("cutoff_motion", ctypes.c_double),
                ("cutoff_appearance", ctypes.c_double)]


class HostBatch(ctypes.Structure):
    _fields_ = [("n_graphs", ctypes.c_int32), ("dim", ctypes.c_int32), ("n", ctypes.c_int64),
                ("graph_ptr", ctypes.c_void_p), ("frame", ctypes.c_void_p), ("pos", ctypes.c_void_p),
                ("conf", ctypes.c_void_p), ("flags", ctypes.c_void_p), ("bbox", ctypes.c_void_p)]


class HostResult(ctypes.Structure):
    _fields_ = [("track_count", ctypes.c_void_p), ("track_ptr", ctypes.c_void_p), ("track_det", ctypes.c_void_p),
                ("total_cost", ctypes.c_void_p), ("n_edges", ctypes.c_int64), ("h2d_bytes", ctypes.c_int64),
                ("d2h_bytes", ctypes.c_int64)]


_vp, _i32, _i64, _sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t
_PD, _PS = ctypes.POINTER(Dets), ctypes.POINTER(CostSpecC)

# name -> (restype, argtypes); every symbol declared in include/mot3d_b200.h
SIGNATURES = {
    "mot3d_version": (ctypes.c_char_p, []),
    "mot3d_last_error": (ctypes.c_char_p, []),
    "mot3d_set_option": (_i32, [ctypes.c_char_p, ctypes.c_double]),
    "mot3d_make_keys": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "mot3d_node_costs": (_i32, [_PD, _PS, _vp, _vp, _vp, _vp]),
    "mot3d_frame_sort": (_i32, [_PD, _vp]),
    "mot3d_edge_count": (_i32, [_PD, _PS, _i32, _vp, _vp]),
    "mot3d_scan_workspace_bytes": (_sz, [_i64]),
    "mot3d_exclusive_scan_i32": (_i32, [_vp, _i64, _vp, _vp, _sz, _vp]),
    "mot3d_edge_fill": (_i32, [_PD, _PS, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mot3d_edge_build_workspace_bytes": (_sz, [_i64]),
    "mot3d_edge_build": (_i32, [_PD, _PS, _i32, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mot3d_edge_appearance": (_i32, [_PD, _PS, _i32, _vp, _vp, _i64, _vp, _vp, _vp]),
    "mot3d_edge_appearance_views": (_i32, [_PD, ctypes.POINTER(Views), _PS, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "mot3d_tracklet_edge_count": (_i32, [ctypes.POINTER(Tracklets), ctypes.POINTER(TrackletSpecC), _vp, _vp]),
    "mot3d_tracklet_edge_fill": (_i32, [ctypes.POINTER(Tracklets), ctypes.POINTER(TrackletSpecC), _vp, _i64, _vp, _vp,
                                        _vp, _vp, _vp]),
    "mot3d_components_workspace_bytes": (_sz, [_i64, _i32]),
    "mot3d_components": (_i32, [_i32, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mot3d_solve_workspace_bytes": (_sz, [_i64, _i32, _i64]),
    "mot3d_solve_batch": (_i32, [_i32, _i64, _vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _vp, _sz, _vp]),
    "mot3d_solve_batch_phase": (_i32, [_i32, _i32, _i32, _i32, _i64, _vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mot3d_order_path_costs": (_i32, [_i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "mot3d_solve_trace_events": (_i32, [_vp, _i32]),
    "mot3d_extract_tracks": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mot3d_track_payload_words": (_i64, [_i64, _i32]),
    "mot3d_pack_tracks": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mot3d_unpack_tracks": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mot3d_peer_buffer_create": (_i32, [_sz, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]),
    "mot3d_peer_buffer_open": (_i32, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "mot3d_peer_buffer_close": (_i32, [_vp]),
    "mot3d_peer_buffer_destroy": (_i32, [_vp]),
    "mot3d_peer_put": (_i32, [_vp, _vp, _sz, _vp, _vp, _i32, _vp]),
    "mot3d_peer_wait": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "mot3d_peer_set_word": (_i32, [_vp, _i32, _vp]),
    "mot3d_peer_wait_word": (_i32, [_vp, _i32, _vp, _vp]),
    "mot3d_merge_close_points": (_i32, [_vp, _i32, _i32, _i32, _vp, ctypes.c_double, _vp, _vp, _vp, _vp]),
    "mot3d_tracklets_from_tracks_workspace_bytes": (_sz, [_i64]),
    "mot3d_tracklets_from_tracks": (_i32, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp,
                                           _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mot3d_tracklet_graph_in_csr": (_i32, [_vp, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                           _vp, _vp]),
    "mot3d_concat_tracklets": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mot3d_online_tracklets": (_i32, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _i64,
                                      _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mot3d_online_concat": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _i64, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mot3d_tracklet_window_hist": (_i32, [_i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                          _vp, _vp, _vp]),
    "mot3d_color_histograms": (_i32, [_vp, _i32, _i32, _i64, _vp, _i32, ctypes.POINTER(ctypes.c_int32), _i32, _i32, _vp, _vp,
                                      _vp]),
    "mot3d_track_batch_workspace_bytes": (_sz, [_i64, _i32, _i64]),
    "mot3d_host_release": (_i32, []),
    "mot3d_track_batch_host": (_i32, [ctypes.POINTER(HostBatch), _PS, ctypes.POINTER(HostResult), _vp, _sz, _i64, _vp]),
}

_lib = None


def lib():
    """The loaded library; raises if it was not built (python __graft_entry__.py build / build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("mot3d_b200: CUDA library %s is missing; build it with `python -c 'import "
                              "__graft_entry__ as g; g.build()'`. There is no CPU fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(code):
    if code != 0:
        raise Mot3dError(code, lib().mot3d_last_error().decode())


def set_option(name, value=None):
    """mot3d_set_option: test hooks / tuning knobs of the library; value=None restores the built-in default."""
    check(lib().mot3d_set_option(name.encode(), float("nan") if value is None else float(value)))


def version():
    return lib().mot3d_version().decode()
