"""Cost specification: the parameters of the built-in cost functions, recovered from the user's callables and
pre-digested for the device (include/mot3d_b200.h: mot3d_cost_spec)."""
import threading

import numpy as np

from . import _lib

SCALE = 10 ** 7  # costs are integers in units of 1e-7, the precision of graph_to_text (reference mot3d/mot3d.py:224)


def quantise_scalar(w):
    """What the reference's text carries for weight w: "{:0.7f}".format(w), as an integer number of 1e-7 units."""
    if not np.isfinite(float(w)):
        raise ValueError("cost %r is not finite: the reference would hand muSSP the text %r" % (w, "{:0.7f}".format(float(w))))
    s = "{:0.7f}".format(float(w))
    neg = s.startswith("-")
    whole, frac = s.lstrip("-").split(".")
    v = int(whole) * SCALE + int(frac)
    return -v if neg else v


def sq_dist_max(max_distance):
    """Largest double s with sqrt(s) <= max_distance, so that the kernel's `s <= sq_dist_max` is exactly the
    reference's `not (dist > max_distance)` (reference mot3d/weight_functions.py:21-23) without a sqrt per pair."""
    m = np.float64(max_distance)
    if not np.isfinite(m):
        return float(np.finfo(np.float64).max)
    if m < 0:
        return -1.0
    s = np.float64(m * m)
    while np.sqrt(np.nextafter(s, np.inf)) <= m:
        s = np.nextafter(s, np.inf)
    while np.sqrt(s) > m:
        s = np.nextafter(s, -np.inf)
    return float(s)


class CostSpec(object):
    """kind: 'detections_2d' | 'detections_3d'. distance = kwargs of weight_distance_detections_2d/_3d,
    confidence = kwargs of weight_confidence_detections (reference mot3d/weight_functions.py:13-17,37-38,43-47)."""

    DIST_DEFAULTS = {
        "detections_2d": dict(sigma_jump=1, sigma_distance=2, sigma_color_histogram=0.3, sigma_box_size=0.3,
                              max_distance=20, use_color_histogram=True, use_bbox=True),
        "detections_3d": dict(sigma_jump=1, sigma_distance=2, sigma_color_histogram=0.3, sigma_box_size=0.3,
                              max_distance=5, use_color_histogram=True, use_bbox=True),
    }

    def __init__(self, kind="detections_2d", distance=None, confidence=None, weight_source_sink=1, max_jump=12):
        if kind not in self.DIST_DEFAULTS:
            raise TypeError("unsupported cost kind %r" % (kind,))
        self.kind = kind
        d = dict(self.DIST_DEFAULTS[kind])
        for k, v in (distance or {}).items():
            if k not in d:
                raise TypeError("unexpected keyword argument %r for weight_distance_%s" % (k, kind))
            d[k] = v
        c = dict(mul=1, bias=0)
        for k, v in (confidence or {}).items():
            if k not in c:
                raise TypeError("unexpected keyword argument %r for weight_confidence_detections" % (k,))
            c[k] = v
        self.distance = d
        self.confidence = c
        self.weight_source_sink = weight_source_sink
        self.max_jump = int(max_jump)

    @property
    def dim(self):
        return 2 if self.kind == "detections_2d" else 3

    @property
    def use_hist(self):
        return bool(self.distance["use_color_histogram"])

    @property
    def use_bbox(self):
        return bool(self.distance["use_bbox"])

    def to_c(self):
        if not (1 <= self.max_jump <= _lib.MAX_JUMP):
            raise ValueError("max_jump must be in [1, %d]" % _lib.MAX_JUMP)
        d = self.distance
        c = _lib.CostSpecC()
        c.max_jump = self.max_jump
        c.use_bbox = int(self.use_bbox)
        c.use_hist = int(self.use_hist)
        c.sq_dist_max = sq_dist_max(d["max_distance"])
        c.sigma_distance_sq = float(d["sigma_distance"] ** 2)
        c.sigma_hist_sq = float(d["sigma_color_histogram"] ** 2)
        c.sigma_bbox_sq = float(d["sigma_box_size"] ** 2)
        c.conf_mul = float(self.confidence["mul"])
        c.conf_bias = float(self.confidence["bias"])
        c.entry_cost = quantise_scalar(self.weight_source_sink)
        c.exit_cost = 0  # reference mot3d/mot3d.py:99
        for k in range(self.max_jump):
            # -np.exp(-(jump-1)*sigma_jump), jump-1 = k (reference mot3d/weight_functions.py:33-34)
            c.neg_exp_jump[k] = float(-np.exp(-(k) * d["sigma_jump"]))
        return c


class TrackletCostSpec(object):
    """kind: 'tracklets_2d' | 'tracklets_3d'; distance = kwargs of weight_distance_tracklets_2d/_3d, confidence =
    kwargs of weight_confidence_tracklets_* (reference mot3d/weight_functions.py:80-84,125-132)."""

    DIST_DEFAULTS = {
        "tracklets_2d": dict(sigma_color_histogram=0.3, sigma_motion=50, alpha=0.7, cutoff_motion=0.1,
                             cutoff_appearance=0.1, max_distance=None, use_color_histogram=True),
        "tracklets_3d": dict(sigma_color_histogram=0.3, sigma_motion=5, alpha=0.7, cutoff_motion=0.1,
                             cutoff_appearance=0.1, max_distance=None, use_color_histogram=True),
    }

    def __init__(self, kind="tracklets_2d", distance=None, confidence=None, weight_source_sink=1, max_jump=12):
        if kind not in self.DIST_DEFAULTS:
            raise TypeError("unsupported cost kind %r" % (kind,))
        self.kind = kind
        d = dict(self.DIST_DEFAULTS[kind])
        for k, v in (distance or {}).items():
            if k not in d:
                raise TypeError("unexpected keyword argument %r for weight_distance_%s" % (k, kind))
            d[k] = v
        c = dict(mul=1, bias=0)
        for k, v in (confidence or {}).items():
            if k not in c:
                raise TypeError("unexpected keyword argument %r for weight_confidence_tracklets" % (k,))
            c[k] = v
        self.distance = d
        self.confidence = c
        self.weight_source_sink = weight_source_sink
        self.max_jump = int(max_jump)

    @property
    def dim(self):
        return 2 if self.kind == "tracklets_2d" else 3

    @property
    def use_hist(self):
        return bool(self.distance["use_color_histogram"])

    def to_c(self):
        d = self.distance
        c = _lib.TrackletSpecC()
        c.max_jump = self.max_jump
        c.use_hist = int(self.use_hist)
        c.check_access_point = 1 if self.kind == "tracklets_2d" else 0
        c.sq_dist_max = -1.0 if d["max_distance"] is None else sq_dist_max(d["max_distance"])
        c.sigma_motion_sq = float(d["sigma_motion"] ** 2)
        c.sigma_hist_sq = float(d["sigma_color_histogram"] ** 2)
        c.alpha = float(d["alpha"])
        c.cutoff_motion = float(d["cutoff_motion"])
        c.cutoff_appearance = float(d["cutoff_appearance"])
        return c


# ---- recovering the spec from opaque callables -------------------------------------------------------------------
# The reference takes arbitrary callables (in practice lambdas closing over kwargs of the built-ins,
# examples/simple_2d/example.py:52-58). build_graph calls the user's callable once on probe objects; the built-in
# cost functions recognise the probes and record which function was reached with which kwargs.
_tls = threading.local()


class Recorded(object):
    def __init__(self, name, kwargs):
        self.name = name
        self.kwargs = kwargs


def is_probe(obj):
    return getattr(obj, "_m3_probe", False)


def record(name, kwargs):
    _tls.last = Recorded(name, dict(kwargs))
    return _tls.last


def take_record():
    r = getattr(_tls, "last", None)
    _tls.last = None
    return r


def probe_callable(fn, *probes):
    """Call fn(*probes) and return what a built-in recorded; TypeError for anything that is not a built-in cost."""
    _tls.last = None
    try:
        fn(*probes)
    except Exception as e:  # a foreign callable touching attributes the probe lacks, etc.
        raise TypeError("mot3d_b200 only accepts the built-in cost functions of mot3d.weight_functions (possibly "
                        "wrapped in a lambda or functools.partial); probing %r failed: %s: %s. There is no CPU "
                        "fallback for arbitrary Python callables." % (fn, type(e).__name__, e))
    r = take_record()
    if r is None:
        raise TypeError("mot3d_b200 only accepts the built-in cost functions of mot3d.weight_functions (possibly "
                        "wrapped in a lambda or functools.partial); %r did not call one. There is no CPU fallback "
                        "for arbitrary Python callables." % (fn,))
    return r
