"""Similarity primitives named as in the reference (mot3d/distance_functions.py). On the hot path the pairwise
ones run inside the CUDA build kernels (csrc/edge_build.cu: bbox_similarity, k_edge_appearance); the functions here
are the per-item helpers users call while preparing detections. `color_histogram` / `color_histograms` (image patch ->
Lab histogram) run on the GPU (csrc/color_hist.cu); OpenCV is not needed."""
import ctypes

import numpy as np

__all__ = ["color_histogram", "color_histograms", "color_histogram_similarity", "bbox_size_similarity",
           "linear_motion"]

_BIN_TABLES = {}


def _bin_of_value(N):
    """Histogram bin of every 8-bit value under np.histogram(bins=N, range=(0, 255)), as a uint16 table."""
    t = _BIN_TABLES.get(N)
    if t is None:
        v = np.arange(256)
        t = np.array([int(np.argmax(np.histogram(v[k:k + 1], bins=N, range=(0, 255))[0])) for k in range(256)],
                     dtype=np.uint16)
        _BIN_TABLES[N] = t
    return t


def color_histograms(image, boxes, channels=(0, 1, 2), N=255, out=None):
    """color_histogram of many patches of one image in a single launch: patch b = image[ymin:ymax, xmin:xmax] for
    boxes[b] = (xmin, ymin, xmax, ymax). image: uint8 [H, W, 3] RGB (numpy or CUDA tensor); boxes: int [n, 4].
    Returns a CUDA float64 tensor [n, len(channels), N] - the `hist` array of a DetectionBatch - without a host
    round trip (reference: one cv2.cvtColor + np.histogram per detection, mot3d/distance_functions.py:10-19)."""
    import torch
    from . import _lib
    if not torch.cuda.is_available():
        raise RuntimeError("mot3d_b200.color_histograms needs a CUDA device; there is no CPU path")
    img = image if torch.is_tensor(image) else torch.from_numpy(np.ascontiguousarray(image))
    if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
        raise ValueError("image must be uint8 [H, W, 3] (RGB), as cv2.cvtColor(patch, cv2.COLOR_RGB2LAB) expects")
    img = img.cuda().contiguous() if not img.is_cuda else img.contiguous()
    dev = img.device
    bx = boxes if torch.is_tensor(boxes) else torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4))
    bx = bx.to(device=dev, dtype=torch.int32).contiguous()
    n, C = int(bx.shape[0]), len(channels)
    if out is None:
        out = torch.empty((n, C, int(N)), dtype=torch.float64, device=dev)
    ch = (ctypes.c_int32 * C)(*[int(c) for c in channels])
    tab = None
    if int(N) != 255:
        tab = torch.from_numpy(_bin_of_value(int(N)).view(np.int16)).to(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mot3d_color_histograms(
            ctypes.c_void_p(img.data_ptr()), int(img.shape[0]), int(img.shape[1]), int(img.stride(0)),
            ctypes.c_void_p(bx.data_ptr()), n, ch, C, int(N), None if tab is None else ctypes.c_void_p(tab.data_ptr()),
            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


def color_histogram(patch, channels=(0, 1, 2), N=255):
    """Mean-subtracted Lab histogram per channel of one uint8 RGB patch (reference mot3d/distance_functions.py:10-19);
    same list of float64 arrays, computed by the GPU kernel (bit-identical: the 8-bit Lab conversion is integer
    arithmetic, tests/test_color_histogram.py)."""
    patch = np.asarray(patch)
    if patch.ndim != 3 or patch.shape[2] != 3 or patch.dtype != np.uint8 or patch.size == 0:
        raise ValueError("patch must be a non-empty uint8 [h, w, 3] RGB array")
    h = color_histograms(patch, np.array([[0, 0, patch.shape[1], patch.shape[0]]]), channels, N)
    h = h[0].cpu().numpy()
    return [h[c] for c in range(len(channels))]


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
