"""merge_close_points on the GPU (reference mot3d/utils/filt.py:6-24): the de-duplication of one frame's detections
that the reference's 3-D example runs before build_graph (examples/soldiers_3d/example.py:29-31). DBSCAN with
min_samples=1 = connected components of "distance <= eps"; every cluster becomes the mean of its members, clusters in
the order of their first member. merge_close_points_batch does all frames of a sequence in one launch
(csrc/filt.cu, one block per frame)."""
import ctypes

import numpy as np
import torch

from .. import _lib

__all__ = ["merge_close_points", "merge_close_points_batch"]


def merge_close_points_batch(point_sets, eps=0.3, return_indexes=False, device=None):
    """point_sets: list of [k_s, dim] point lists (one per frame). Returns one result per set: the merged points
    (array [clusters, dim]) or, with return_indexes, the member indexes of every cluster."""
    L = _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    sizes = [len(p) for p in point_sets]
    n = int(sum(sizes))
    if n == 0:
        return [p for p in point_sets]
    dim = int(np.asarray(next(p for p in point_sets if len(p))).reshape(len(next(p for p in point_sets if len(p))), -1).shape[1])
    flat = np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, dim) for p in point_sets if len(p)])
    set_ptr = np.cumsum([0] + sizes).astype(np.int32)
    d_ptr = torch.from_numpy(set_ptr).to(dev)
    d_pts = torch.from_numpy(np.ascontiguousarray(flat)).to(dev)
    d_label = torch.empty(n, dtype=torch.int32, device=dev)
    d_count = torch.empty(len(sizes), dtype=torch.int32, device=dev)
    d_out = torch.empty((n, dim), dtype=torch.float64, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(L.mot3d_merge_close_points(p(d_ptr), len(sizes), int(max(sizes)), dim, p(d_pts), float(eps), p(d_label),
                                          p(d_count), p(d_out),
                                          ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    label, count, out = d_label.cpu().numpy(), d_count.cpu().numpy(), d_out.cpu().numpy()
    res = []
    for s, k in enumerate(sizes):
        a = int(set_ptr[s])
        if k == 0:
            res.append(point_sets[s])
        elif return_indexes:
            res.append([np.flatnonzero(label[a:a + k] == c).tolist() for c in range(int(count[s]))])
        else:
            res.append(out[a:a + int(count[s])].copy())
    return res


def merge_close_points(points, eps=0.3, return_indexes=False):
    """Same call as the reference's (mot3d/utils/filt.py:6): one set of points."""
    if len(points) == 0:
        return points
    return merge_close_points_batch([points], eps=eps, return_indexes=return_indexes)[0]
