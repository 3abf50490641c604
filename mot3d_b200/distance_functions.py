"""Similarity primitives named as in the reference (mot3d/distance_functions.py). On the hot path the pairwise
ones run inside the CUDA build kernels (csrc/edge_build.cu: bbox_similarity, k_edge_appearance); the functions here
are the per-item helpers users call while preparing detections. `color_histogram` (image -> LAB histogram) is input
preparation, outside the hot path."""
import numpy as np

__all__ = ["color_histogram", "color_histogram_similarity", "bbox_size_similarity", "linear_motion"]


def color_histogram(patch, channels=(0, 1, 2), N=255):
    """Mean-subtracted LAB histogram per channel (reference mot3d/distance_functions.py:10-19). Needs OpenCV."""
    import cv2
    lab = cv2.cvtColor(patch, cv2.COLOR_RGB2LAB)
    out = []
    for c in channels:
        h = np.histogram(lab[:, :, c].ravel(), bins=N, range=(0, 255))[0]
        out.append(h - h.mean())
    return out


def color_histogram_similarity(H1, H2, weights=None):
    """Mean over channels of the normalised cross-correlation (reference mot3d/distance_functions.py:21-34)."""
    C = len(H1)
    if weights is None:
        weights = (1,) * C
    acc = 0
    for c in range(C):
        a, b = np.asarray(H1[c]), np.asarray(H2[c])
        acc += weights[c] * (np.sum(a * b) / np.sqrt(np.sum(a ** 2) * np.sum(b ** 2)))
    return acc / C


def bbox_size_similarity(bbox1, bbox2):
    """1 - mean relative difference of height and width (reference mot3d/distance_functions.py:36-45)."""
    h1, w1 = bbox1[3] - bbox1[1], bbox1[2] - bbox1[0]
    h2, w2 = bbox2[3] - bbox2[1], bbox2[2] - bbox2[0]
    return 1 - (np.abs(h1 - h2) / (np.maximum(h1, h2) + 1e-12) + np.abs(w1 - w2) / (np.maximum(w1, w2) + 1e-12)) / 2


def linear_motion(idxs1, pts1, idxs2, pts2):
    """Two-way deviation of degree-1 extrapolations between two point chains (reference
    mot3d/distance_functions.py:47-63): fit a line through one chain, evaluate at the other's times, average the
    L2 errors; sum of both directions."""
    assert idxs2[0] > idxs1[-1], "idxs2[0]={} idxs1[-1]={} tracklet2 cannot be more recent than tracklet1!".format(
        idxs2[0], idxs1[-1])

    def one_way(ta, pa, tb, pb):
        coef = np.polyfit(ta, pa, 1)                       # [2, dim]
        pred = np.outer(np.asarray(tb, dtype=float), coef[0]) + coef[1]
        return np.linalg.norm(pred - np.asarray(pb), axis=1).mean()

    return one_way(idxs1, pts1, idxs2, pts2) + one_way(idxs2, pts2, idxs1, pts1)
