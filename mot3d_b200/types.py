"""Detection containers with the reference's attribute names (reference mot3d/types.py:11-45 for detections,
:47-184 for tracklets). They are plain records; build_graph packs them into structure-of-arrays tensors."""
import numpy as np

__all__ = ["Detection", "Detection2D", "Detection3D", "DetectionTracklet", "DetectionTracklet2D",
           "DetectionTracklet3D", "TrajectoryView"]


class _GraphItem(object):
    """What build_graph reads and writes on every item: confidence, id, node back-references, terminal flags."""

    def _init_item(self, confidence, id):
        self.confidence = confidence
        self.id = id
        self.pre_node = None      # set by build_graph (reference mot3d/mot3d.py:89)
        self.post_node = None     # (reference mot3d/mot3d.py:97)
        self.weight_source = None
        self.weight_sink = None
        self.entry_points = True  # S -> pre arc allowed
        self.exit_point = True    # post -> T arc allowed


class Detection(_GraphItem):
    """A detection at frame `index`. diff_index(other) = other.index - self.index is the time gap that the
    max_jump window is applied to (reference mot3d/types.py:27-28)."""

    def __init__(self, index, confidence=0.5, id=None):
        self.index = index
        self._init_item(confidence, id)

    def diff_index(self, other_detection):
        return other_detection.index - self.index


class Detection2D(Detection):
    def __init__(self, index, position=None, confidence=0.5, id=None, view=None, color_histogram=None, bbox=None):
        Detection.__init__(self, index, confidence, id)
        self.position = position
        self.view = view
        self.color_histogram = color_histogram  # list of C arrays of B bins, mean-subtracted
        self.bbox = bbox                        # (xmin, ymin, xmax, ymax)


class Detection3D(Detection):
    def __init__(self, index, position=None, confidence=0.5, id=None, detections_2d={}):
        Detection.__init__(self, index, confidence, id)
        self.position = position
        self.detections_2d = detections_2d      # {view: Detection2D}


class DetectionTracklet(_GraphItem):
    """A chain of detections treated as one node. head = most recent end, tail = oldest end; the time gap between
    two tracklets is other.tail.index - self.head.index (reference mot3d/types.py:47-68)."""

    def __init__(self, tracklet, confidence=0.5, id=None):
        self.tracklet = tracklet
        self._init_item(confidence, id)
        self.head = tracklet[-1]
        self.tail = tracklet[0]

    def diff_index(self, other_tracklet):
        return other_tracklet.tail.index - self.head.index


def _require(tracklet, cls):
    for d in tracklet:
        if not isinstance(d, cls):
            raise ValueError("The tracklet must be composed of objects of type {} only! ({})".format(cls.__name__, type(d)))


def _end_features_2d(dets):
    idx = [d.index for d in dets]
    pts = [d.position for d in dets]
    hists = [d.color_histogram for d in dets if d.color_histogram is not None]
    boxes = [d.bbox for d in dets if d.bbox is not None]
    hist = np.median(hists, axis=0) if len(hists) else None
    box = np.median(boxes, axis=0) if len(boxes) else None
    return idx, pts, hist, box


class DetectionTracklet2D(DetectionTracklet):
    """2D tracklet; each end summarises the detections within `window` frames of it: their indexes/positions (for the
    linear motion model) and the median histogram/bbox (reference mot3d/types.py:94-131)."""

    def __init__(self, tracklet, confidence=0.5, id=None, window=None, view=None,
                 head_access_point=False, tail_access_point=False):
        DetectionTracklet.__init__(self, tracklet, confidence, id)
        _require(tracklet, Detection2D)
        self.window = window
        self.view = view
        if window is None:
            head_part = tail_part = tracklet
        else:
            last, first = tracklet[-1].index, tracklet[0].index
            head_part = [d for d in tracklet if d.index > last - window]
            tail_part = [d for d in tracklet if d.index < first + window]
        idx, pts, hist, box = _end_features_2d(head_part)
        self.head = Detection2D(idx[-1], pts[-1], view=view, color_histogram=hist, bbox=box)
        self.head.indexes, self.head.positions, self.head.access_point = idx, pts, head_access_point
        if window is not None:
            idx, pts, hist, box = _end_features_2d(tail_part)
        self.tail = Detection2D(idx[0], pts[0], view=view, color_histogram=hist, bbox=box)
        self.tail.indexes, self.tail.positions, self.tail.access_point = idx, pts, tail_access_point


class DetectionTracklet3D(DetectionTracklet):
    """3D tracklet; also regroups the per-view 2D detections into one DetectionTracklet2D per view
    (reference mot3d/types.py:142-184)."""

    def __init__(self, tracklet, confidence=0.5, id=None, window=None,
                 head_access_point=False, tail_access_point=False):
        DetectionTracklet.__init__(self, tracklet, confidence, id)
        _require(tracklet, Detection3D)
        self.window = window
        per_view = {}
        for d in tracklet:
            for view, d2 in d.detections_2d.items():
                per_view.setdefault(view, []).append(d2)
        self.tracklets_2d = {view: DetectionTracklet2D(chain, view=view, window=window)
                             for view, chain in per_view.items()}
        if window is None:
            head_part = tail_part = tracklet
        else:
            last, first = tracklet[-1].index, tracklet[0].index
            head_part = [d for d in tracklet if d.index > last - window]
            tail_part = [d for d in tracklet if d.index < first + window]
        idx, pts = [d.index for d in head_part], [d.position for d in head_part]
        self.head = Detection3D(idx[-1], pts[-1])
        self.head.indexes, self.head.positions, self.head.access_point = idx, pts, head_access_point
        if window is not None:
            idx, pts = [d.index for d in tail_part], [d.position for d in tail_part]
        self.tail = Detection3D(idx[0], pts[0])
        self.tail.indexes, self.tail.positions, self.tail.access_point = idx, pts, tail_access_point


class TrajectoryView(object):
    """Placeholder pointing at a piece of a trajectory stored elsewhere (reference mot3d/types.py:191-193)."""

    def __init__(self, id_view=None):
        self.id_view = id_view
