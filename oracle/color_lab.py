"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's color_histogram
(/root/reference/mot3d/distance_functions.py:10-19) without OpenCV: 8-bit RGB -> Lab by the integer algorithm of
cv2.cvtColor(uint8, COLOR_RGB2LAB) (OpenCV 4.x color_lab.cpp, RGB2Lab_b; OpenCV is a third-party dependency of the
reference, not vendored; checked here against cv2 4.13.0), then np.histogram per channel minus its mean.
The tables are read from the very header the CUDA kernel compiles (mot3d_b200/csrc/lab_tables.h), so pinning this
restatement pins the kernel's constants: tests/test_color_histogram.py compares it with cv2 on all 2^24 colours and
with the golden fixture made by the reference function (tests/golden/color_histogram.npz).
Only tests/ may import this module."""
import os
import re

import numpy as np

_HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mot3d_b200", "csrc", "lab_tables.h")
_T = {}


def tables():
    if not _T:
        src = open(_HDR).read()
        for name in ("LAB_SHIFT", "LAB_SHIFT2", "LAB_L_SCALE", "LAB_L_SHIFT"):
            _T[name] = int(re.search(r"#define %s \(?(-?\d+)\)?" % name, src).group(1))
        for name in ("LAB_COEF", "LAB_GAMMA_TAB", "LAB_CBRT_TAB"):
            body = re.search(r"%s\[\d+\] = \{([^}]*)\}" % name, src).group(1)
            _T[name] = np.array([int(x) for x in body.replace("\n", " ").split(",")], dtype=np.int64)
    return _T


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def rgb_to_lab_u8(rgb):
    """rgb: uint8 [..., 3] -> uint8 [..., 3] (L, a, b) as cv2.cvtColor(rgb, cv2.COLOR_RGB2LAB)."""
    t = tables()
    g, cb, C = t["LAB_GAMMA_TAB"], t["LAB_CBRT_TAB"], t["LAB_COEF"]
    s1, s2 = t["LAB_SHIFT"], t["LAB_SHIFT2"]
    rgb = np.asarray(rgb)
    R, G, B = g[rgb[..., 0]], g[rgb[..., 1]], g[rgb[..., 2]]
    fX = cb[_descale(R * C[0] + G * C[1] + B * C[2], s1)]
    fY = cb[_descale(R * C[3] + G * C[4] + B * C[5], s1)]
    fZ = cb[_descale(R * C[6] + G * C[7] + B * C[8], s1)]
    L = _descale(t["LAB_L_SCALE"] * fY + t["LAB_L_SHIFT"], s2)
    a = _descale(500 * (fX - fY) + 128 * (1 << s2), s2)
    b = _descale(200 * (fY - fZ) + 128 * (1 << s2), s2)
    return np.clip(np.stack([L, a, b], axis=-1), 0, 255).astype(np.uint8)


def color_histogram(patch, channels=(0, 1, 2), N=255):
    lab = rgb_to_lab_u8(patch)
    out = []
    for c in channels:
        h = np.histogram(lab[:, :, c].ravel(), bins=N, range=(0, 255))[0]
        out.append(h - h.mean())
    return out
