"""Trajectory glue between the detection pass and the tracklet pass (reference mot3d/utils/trajectory.py:11-12,
216-248): cutting solved trajectories into tracklets, dropping the short ones, joining linked tracklets back.
Plain list manipulation on the caller's objects (host side: a few hundred items per window); the rest of the
reference's utils (interpolation, smoothing, plotting) is post-processing outside the hot path."""

__all__ = ["concat_tracklets", "split_trajectory_modulo", "remove_short_trajectories"]


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
