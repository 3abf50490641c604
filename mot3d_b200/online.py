"""Online sliding-window tracking: the caller that turns per-window solves into a tracker.

Mirror of the reference's single-view loop - Tracker2D.update (projects/PETS2009S2L1/singleview_tracking.py:76-134),
compute_trajectories / concatenate_tracklets (projects/PETS2009S2L1/tracking_routines.py:16-52) and Scene
(projects/PETS2009S2L1/scene.py:8-52) - with the same names, arguments and bookkeeping. Per batch of frames:
detections -> build_graph + solve_graph (CUDA) -> tracklets cut every 10 frames, short ones dropped -> tracklet graph
over the active trajectories + the new tracklets -> build_graph + solve_graph (CUDA) -> concatenated, hole-filled
trajectories written back to the scene. Everything outside the two build/solve pairs is list bookkeeping on the
caller's objects, as in the reference."""
import itertools
import pickle

from . import mot3d as _m
from .types import Detection2D, DetectionTracklet2D
from .utils import trajectory as utils_traj
from .weight_functions import (weight_distance_detections_2d, weight_confidence_detections_2d,
                               weight_distance_tracklets_2d, weight_confidence_tracklets_2d)

__all__ = ["Scene", "Tracker2D", "compute_trajectories", "concatenate_tracklets"]


def compute_trajectories(detections, weight_source_sink, weight_distance, weight_confidence, max_jump=4,
                         verbose=False):
    """detections: list (one entry per frame of the batch) of lists of detections
    (reference tracking_routines.py:16-34)."""
    detections = list(itertools.chain(*detections))
    g = _m.build_graph(detections, weight_source_sink=weight_source_sink, max_jump=max_jump, verbose=verbose,
                       weight_distance=weight_distance, weight_confidence=weight_confidence)
    return [] if g is None else _m.solve_graph(g, verbose=verbose)


def concatenate_tracklets(detections_tracklets, weight_source_sink, weight_confidence_tracklets,
                          weight_distance_tracklets, max_jump=100, verbose=False):
    """(reference tracking_routines.py:36-52)"""
    g = _m.build_graph(detections_tracklets, weight_source_sink=weight_source_sink, max_jump=max_jump,
                       verbose=verbose, weight_confidence=weight_confidence_tracklets,
                       weight_distance=weight_distance_tracklets)
    return [] if g is None else _m.solve_graph(g, verbose=verbose)


class Scene(object):
    """Active / completed trajectories and the tracklets seen so far (reference scene.py:8-52)."""

    def __init__(self, completed_after=30):
        self.completed_after = completed_after
        self.id_last = 0
        self.active = {}
        self.completed = {}
        self.id_last_tracklet = 0
        self.tracklets = {}

    def store_tracklet(self, tracklet):
        self.tracklets[self.id_last_tracklet] = tracklet
        self.id_last_tracklet += 1

    def update(self, id, trajectory):
        # update an existing trajectory or create a new one
        if id in self.active:
            self.active[id] = trajectory
        else:
            self.active[self.id_last] = trajectory
            self.id_last += 1

    def update_completed(self, time_index):
        # trajectories that have not been updated for a while are completed
        for id, trajectory in list(self.active.items()):
            if (trajectory[-1].index + self.completed_after) < time_index:
                self.completed[id] = self.active[id]
                del self.active[id]

    def save(self, filename):
        with open(filename, "wb") as f:
            pickle.dump({"active": self.active, "completed": self.completed, "trackelts": self.tracklets}, f)

    def load(self, filename):
        with open(filename, "rb") as f:
            data = pickle.load(f)
        self.active = data["active"]
        self.completed = data["completed"]
        self.tracklets = data["trackelts"]
        self.id_last = len(self.active) + len(self.completed)
        self.id_last_tracklet = len(self.tracklets)


class Tracker2D(object):
    """config = the `tracking` section of the reference's YAML
    (projects/PETS2009S2L1/config/config_singleview_PETS2009S2L1_View_001.yaml:18-55): batch_size, detections
    {weight_source_sink, max_jump, verbose, weight_distance{..}, weight_confidence{..}}, tracklets {weight_source_sink,
    max_jump, length_endpoints, verbose, weight_distance{..}, weight_confidence{..}}. valid_region: object with
    is_inside(point) or None (reference singleview_tracking.py:28-38)."""

    def __init__(self, scene, valid_region, config, image_size=None):
        self.scene = scene
        self.valid_region = valid_region
        self.__c__ = config
        self.batch_size = config["batch_size"]
        self.image_size = image_size
        cfgd = config["detections"]
        self.__Detection__ = Detection2D
        wd_d = lambda d1, d2: weight_distance_detections_2d(d1, d2, **cfgd["weight_distance"])
        wc_d = lambda d: weight_confidence_detections_2d(d, **cfgd["weight_confidence"])
        self.__compute_trajs__ = lambda detections: compute_trajectories(
            detections, weight_source_sink=cfgd["weight_source_sink"], weight_distance=wd_d, weight_confidence=wc_d,
            max_jump=cfgd["max_jump"], verbose=cfgd.get("verbose", False))
        cfgt = config["tracklets"]
        self.__DetectionTracklet__ = lambda id, track, **kwargs: DetectionTracklet2D(
            track, id=id, window=cfgt["length_endpoints"], **kwargs)
        wd_t = lambda t1, t2: weight_distance_tracklets_2d(t1, t2, **cfgt["weight_distance"])
        wc_t = lambda t: weight_confidence_tracklets_2d(t, **cfgt["weight_confidence"])
        self.__cat_tracklets__ = lambda tracklets: concatenate_tracklets(
            tracklets, weight_source_sink=cfgt["weight_source_sink"], weight_distance_tracklets=wd_t,
            weight_confidence_tracklets=wc_t, max_jump=cfgt["max_jump"], verbose=cfgt.get("verbose", False))

    def update(self, time_index, detections):
        """One batch (reference singleview_tracking.py:76-134). detections: per frame, a list of Detection2D."""
        if self.valid_region is not None:
            detections = [[d for d in det_frame if self.valid_region.is_inside(d.position)]
                          for det_frame in detections]
        # tracklets of this batch
        tracklets = self.__compute_trajs__(detections)
        tracklet_splits = []
        for tracklet in tracklets:
            tracklet_splits += utils_traj.split_trajectory_modulo(tracklet, length=10)
        tracklet_splits = utils_traj.remove_short_trajectories(tracklet_splits, th_length=2)
        # concatenate them to the active trajectories
        detections_tracklets = [self.__DetectionTracklet__(id, track) for id, track in self.scene.active.items()]
        n_skipped = 0
        for tracklet in tracklet_splits:
            if self.valid_region is not None:
                res1 = self.valid_region.is_inside(tracklet[0].position)   # tail
                res2 = self.valid_region.is_inside(tracklet[-1].position)  # head
            else:
                res1, res2 = True, True
            if res1 and res2:
                detections_tracklets.append(self.__DetectionTracklet__(None, tracklet, tail_access_point=not res1,
                                                                       head_access_point=not res2))
                self.scene.store_tracklet(tracklet)
            else:
                n_skipped += 1  # both endpoints in the access region
        det_trajectories = self.__cat_tracklets__(detections_tracklets)
        for det_trajectory in det_trajectories:
            # an id that is not None: new chunk(s) appended to an existing trajectory
            id = det_trajectory[0].id
            trajectory = utils_traj.concat_tracklets([dt.tracklet for dt in det_trajectory])
            trajectory = utils_traj.interpolate_trajectory(trajectory, attr_names=["position", "bbox"])
            self.scene.update(id, trajectory)
        self.scene.update_completed(time_index)
        return n_skipped
