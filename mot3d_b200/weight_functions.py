"""Built-in arc costs, with the reference's names and keyword arguments (reference mot3d/weight_functions.py).

In this package they are *cost descriptors*: the arithmetic runs in the CUDA kernels (csrc/edge_build.cu).
build_graph calls the user's callable once with probe objects; a built-in reached with probes records its name and
keyword arguments (spec.record) so that lambdas / functools.partial wrappers keep working unmodified.  Called on
real detections they evaluate that single pair on the GPU through the same kernels; there is no CPU arithmetic.

    w_ij = -exp(-(jump-1)*sigma_jump) * exp(-dist^2/sigma_distance^2) [* exp(-(1-NCC)^2/sigma_ch^2)]
                                                                       [* exp(-(1-bss)^2/sigma_bs^2)]
           or None when dist > max_distance                      (reference mot3d/weight_functions.py:13-35, 43-75)
    w_i  = -(confidence*mul + bias)                              (reference mot3d/weight_functions.py:37-38)
"""
from . import spec as _spec

__all__ = ["weight_distance_detections_2d", "weight_confidence_detections", "weight_confidence_detections_2d",
           "weight_distance_detections_3d", "weight_confidence_detections_3d", "weight_distance_tracklets_2d",
           "weight_confidence_tracklets_2d", "weight_distance_tracklets_3d", "weight_confidence_tracklets_3d",
           "weight_distance_detections_auto", "weight_confidence_detections_auto",
           "weight_distance_tracklets_auto", "weight_confidence_tracklets_auto", "project_points",
           "weight_distance_tracklet3d_tracklet2d", "weight_distance_tracklet2d_tracklet3d"]


def _pair_on_gpu(kind, d1, d2, kwargs):
    from .mot3d import _evaluate_pair
    return _evaluate_pair(kind, d1, d2, kwargs)


def weight_distance_detections_2d(d1, d2, sigma_jump=1, sigma_distance=2, sigma_color_histogram=0.3,
                                  sigma_box_size=0.3, max_distance=20, use_color_histogram=True, use_bbox=True):
    kw = dict(sigma_jump=sigma_jump, sigma_distance=sigma_distance, sigma_color_histogram=sigma_color_histogram,
              sigma_box_size=sigma_box_size, max_distance=max_distance, use_color_histogram=use_color_histogram,
              use_bbox=use_bbox)
    if _spec.is_probe(d1):
        return _spec.record("detections_2d", kw)
    return _pair_on_gpu("detections_2d", d1, d2, kw)


def weight_distance_detections_3d(d1, d2, sigma_jump=1, sigma_distance=2, sigma_color_histogram=0.3,
                                  sigma_box_size=0.3, max_distance=5, use_color_histogram=True, use_bbox=True):
    kw = dict(sigma_jump=sigma_jump, sigma_distance=sigma_distance, sigma_color_histogram=sigma_color_histogram,
              sigma_box_size=sigma_box_size, max_distance=max_distance, use_color_histogram=use_color_histogram,
              use_bbox=use_bbox)
    if _spec.is_probe(d1):
        return _spec.record("detections_3d", kw)
    return _pair_on_gpu("detections_3d", d1, d2, kw)


def weight_confidence_detections(detection, mul=1, bias=0):
    if _spec.is_probe(detection):
        return _spec.record("confidence", dict(mul=mul, bias=bias))
    # a single real item: the reference's one-line value (mot3d/weight_functions.py:37-38); build_graph itself never
    # comes here (k_node_costs computes the same expression for all detections at once)
    return -(detection.confidence * mul + bias)


def weight_confidence_detections_2d(detection, **kwargs):
    return weight_confidence_detections(detection, **kwargs)


def weight_confidence_detections_3d(detection, **kwargs):
    return weight_confidence_detections(detection, **kwargs)


def _tracklet_pair_on_gpu(kind, t1, t2, kwargs):
    from .mot3d import _evaluate_tracklet_pair
    return _evaluate_tracklet_pair(kind, t1, t2, kwargs)


def weight_distance_tracklets_2d(t1, t2, sigma_color_histogram=0.3, sigma_motion=50, alpha=0.7, cutoff_motion=0.1,
                                 cutoff_appearance=0.1, max_distance=None, use_color_histogram=True, debug=False):
    kw = dict(sigma_color_histogram=sigma_color_histogram, sigma_motion=sigma_motion, alpha=alpha,
              cutoff_motion=cutoff_motion, cutoff_appearance=cutoff_appearance, max_distance=max_distance,
              use_color_histogram=use_color_histogram)
    if _spec.is_probe(t1):
        return _spec.record("tracklets_2d", kw)
    return _tracklet_pair_on_gpu("tracklets_2d", t1, t2, kw)


def weight_distance_tracklets_3d(t1, t2, sigma_color_histogram=0.3, sigma_motion=5, alpha=0.7, cutoff_motion=0.1,
                                 cutoff_appearance=0.1, max_distance=None, use_color_histogram=True):
    kw = dict(sigma_color_histogram=sigma_color_histogram, sigma_motion=sigma_motion, alpha=alpha,
              cutoff_motion=cutoff_motion, cutoff_appearance=cutoff_appearance, max_distance=max_distance,
              use_color_histogram=use_color_histogram)
    if _spec.is_probe(t1):
        return _spec.record("tracklets_3d", kw)
    return _tracklet_pair_on_gpu("tracklets_3d", t1, t2, kw)


def weight_confidence_tracklets_2d(tracklet, **kwargs):
    return weight_confidence_detections(tracklet, **kwargs)


def weight_confidence_tracklets_3d(tracklet, **kwargs):
    return weight_confidence_detections(tracklet, **kwargs)


def _is(obj, name):
    return getattr(obj, "_m3_kind", None) == name or type(obj).__name__ == name or any(
        c.__name__ == name for c in type(obj).__mro__)


def weight_distance_detections_auto(d1, d2, config2d={}, config3d={}):
    """Type dispatch of the reference (mot3d/weight_functions.py:237-256)."""
    if _is(d1, "Detection2D") and _is(d2, "Detection2D"):
        return weight_distance_detections_2d(d1, d2, **config2d)
    if _is(d1, "Detection3D") and _is(d2, "Detection3D"):
        return weight_distance_detections_3d(d1, d2, **config3d)
    raise RuntimeError("Combination not possible or not implemented [{},{}].".format(type(d1), type(d2)))


def weight_confidence_detections_auto(d, config2d={}, config3d={}):
    if _is(d, "Detection2D"):
        return weight_confidence_detections_3d(d, **config2d)
    if _is(d, "Detection3D"):
        return weight_confidence_detections_3d(d, **config3d)
    raise RuntimeError("Unable to pick the pick the correct weight confidence for detections [{}].".format(type(d)))


_MIXED = ("mixed 2-D / 3-D tracklet costs are not provided: the reference's own implementation cannot run "
          "(mot3d/weight_functions.py:174 passes np.float32 as the `order` argument of np.reshape, a TypeError on "
          "every input), so there is no behaviour to reproduce")


def project_points(points, calib):
    raise NotImplementedError(_MIXED)


def weight_distance_tracklet3d_tracklet2d(t1, t2, calib={}, **kwargs):
    raise NotImplementedError(_MIXED)


def weight_distance_tracklet2d_tracklet3d(t1, t2, calib={}, **kwargs):
    raise NotImplementedError(_MIXED)


def weight_distance_tracklets_auto(t1, t2, config2d={}, config3d={}, calib={}):
    """Type dispatch of the reference (mot3d/weight_functions.py:201-221)."""
    is_t1_2d, is_t2_2d = _is(t1, "DetectionTracklet2D"), _is(t2, "DetectionTracklet2D")
    is_t1_3d, is_t2_3d = _is(t1, "DetectionTracklet3D"), _is(t2, "DetectionTracklet3D")
    if is_t1_2d and is_t2_2d:
        return weight_distance_tracklets_2d(t1, t2, **config2d)
    if is_t1_3d and is_t2_3d:
        return weight_distance_tracklets_3d(t1, t2, **config3d)
    if (is_t1_3d and is_t2_2d) or (is_t1_2d and is_t2_3d):
        raise NotImplementedError(_MIXED)
    raise RuntimeError("Unable to pick the pick the correct weight distance for tracklets [{},{}].".format(type(t1),
                                                                                                           type(t2)))


def weight_confidence_tracklets_auto(t, config2d={}, config3d={}):
    if _is(t, "DetectionTracklet2D"):
        return weight_confidence_tracklets_2d(t, **config2d)
    if _is(t, "DetectionTracklet3D"):
        return weight_confidence_tracklets_3d(t, **config3d)
    raise RuntimeError("Unable to pick the pick the correct weight confidence for tracklets [{}].".format(type(t)))
