"""Trajectory glue between the detection pass and the tracklet pass (reference mot3d/utils/trajectory.py:11-12,
216-248): cutting solved trajectories into tracklets, dropping the short ones, joining linked tracklets back.
Plain list manipulation on the caller's objects (host side: a few hundred items per window), plus the hole filling
the online loop applies to every linked trajectory (reference :14-69); the rest of the reference's utils (smoothing,
plotting) is post-processing outside the hot path."""
import numpy as np

__all__ = ["concat_tracklets", "split_trajectory_modulo", "remove_short_trajectories", "interpolate_trajectory"]


def concat_tracklets(tracklets):
    """One trajectory from a chain of tracklets (reference mot3d/utils/trajectory.py:11-12)."""
    out = []
    for t in tracklets:
        out.extend(t)
    return out


def split_trajectory_modulo(trajectory, length=10):
    """Cut a trajectory wherever `index % length` wraps around, i.e. into pieces that each stay inside one block of
    `length` frames (reference mot3d/utils/trajectory.py:216-241)."""
    pieces = []
    current = []
    last = -1
    for d in trajectory:
        m = d.index % length
        if current and m < last:
            pieces.append(current)
            current = []
        current.append(d)
        last = m
    if current:
        pieces.append(current)
    return pieces


def remove_short_trajectories(trajectories, th_length=10, verbose=False):
    """Keep the trajectories with more than th_length detections (reference mot3d/utils/trajectory.py:243-248)."""
    return [t for t in trajectories if len(t) > th_length]


def interpolate_trajectory(trajectory, attr_names=['position']):
    """Fill the frames a trajectory skips with linearly interpolated detections of the same type; existing detections
    are kept as they are (reference mot3d/utils/trajectory.py:14-69; scipy's interp1d there evaluates
    slope * (t - t1) + v1 with slope = (v2 - v1) / (t2 - t1), restated here)."""
    cls = type(trajectory[0])
    out = [trajectory[0]]
    for curr in trajectory[1:]:
        past = out[-1]
        if curr.index - past.index != 1:
            new_indexes = np.arange(past.index + 1, curr.index)
            new_attrs = {}
            for name in attr_names:
                x1, x2 = getattr(past, name, None), getattr(curr, name, None)
                if x1 is None or x2 is None:
                    raise ValueError("Detection does not have attribute {}!".format(name))
                x1, x2 = np.asarray(x1), np.asarray(x2)
                v1, v2 = np.ravel(x1).astype(np.float64), np.ravel(x2).astype(np.float64)
                slope = (v2 - v1) / float(curr.index - past.index)
                vals = slope[None, :] * (new_indexes - past.index)[:, None].astype(np.float64) + v1[None, :]
                new_attrs[name] = vals.reshape((-1,) + x1.shape)
            for i, index in enumerate(new_indexes):
                out.append(cls(index=index, **{name: attr[i] for name, attr in new_attrs.items()}))
        out.append(curr)
    return out
