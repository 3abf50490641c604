"""color_histogram (reference mot3d/distance_functions.py:10-19): oracle vs the golden fixture made by
the reference function and vs OpenCV on the whole 8-bit colour cube (CPU); CUDA kernel vs both (GPU). Integer work:
bit-exact, including the float64 "count - mean" values."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
VARIANTS = {"default": {}, "ch12_n64": dict(channels=(1, 2), N=64), "ch0_n100": dict(channels=(0,), N=100),
            "ch210_n300": dict(channels=(2, 1, 0), N=300)}


def _golden():
    return np.load(os.path.join(HERE, "golden", "color_histogram.npz"))


def test_oracle_matches_reference_fixture():
    from oracle import color_lab
    d = _golden()
    img = d["image"]
    for tag, kw in VARIANTS.items():
        for b, (xmin, ymin, xmax, ymax) in enumerate(d["boxes"]):
            got = np.array(color_lab.color_histogram(img[ymin:ymax, xmin:xmax], **kw))
            assert (got == d["hist_" + tag][b]).all(), (tag, b)


def test_lab_tables_against_opencv_on_every_colour():
    """The tables the kernel compiles (csrc/lab_tables.h) reproduce cv2.cvtColor(uint8, COLOR_RGB2LAB) on all 2^24
    colours."""
    cv2 = pytest.importorskip("cv2")
    from oracle import color_lab
    v = np.arange(256, dtype=np.uint8)
    for r0 in range(0, 256, 8):
        img = np.empty((8, 256, 256, 3), np.uint8)
        img[..., 0] = np.arange(r0, r0 + 8, dtype=np.uint8)[:, None, None]
        img[..., 1] = v[None, :, None]
        img[..., 2] = v[None, None, :]
        img = img.reshape(8 * 256, 256, 3)
        assert (cv2.cvtColor(img, cv2.COLOR_RGB2LAB) == color_lab.rgb_to_lab_u8(img)).all(), r0


def test_bin_table_is_numpy_histogram():
    from mot3d_b200.distance_functions import _bin_of_value
    for N in (255, 64, 100, 300, 1, 7):
        t = _bin_of_value(N)
        h = np.zeros(N, dtype=np.int64)
        vals = np.arange(256).repeat(3)
        np.add.at(h, t[vals], 1)
        assert (h == np.histogram(vals, bins=N, range=(0, 255))[0]).all()


@pytest.mark.gpu
def test_kernel_matches_reference_fixture():
    import torch
    import mot3d_b200 as mot3d
    d = _golden()
    img, boxes = d["image"], d["boxes"]
    for tag, kw in VARIANTS.items():
        got = mot3d.color_histograms(img, boxes, **kw)
        assert got.dtype == torch.float64 and got.is_cuda
        assert (got.cpu().numpy() == d["hist_" + tag]).all(), tag
    # the reference's per-patch signature
    xmin, ymin, xmax, ymax = boxes[0]
    one = mot3d.color_histogram(img[ymin:ymax, xmin:xmax])
    assert len(one) == 3 and all((one[c] == d["hist_default"][0][c]).all() for c in range(3))
    with pytest.raises(ValueError):
        mot3d.color_histogram(img[0:0, 0:0])
    # the histograms feed the appearance term directly: same similarity as on the reference's lists
    got = mot3d.color_histograms(img, boxes).cpu().numpy()
    s_ref = mot3d.color_histogram_similarity(list(d["hist_default"][1]), list(d["hist_default"][6]))
    s_got = mot3d.color_histogram_similarity(list(got[1]), list(got[6]))
    assert np.isfinite(s_ref) and s_got == s_ref


@pytest.mark.gpu
def test_kernel_matches_oracle_on_a_large_frame():
    import mot3d_b200 as mot3d
    from oracle import color_lab
    rng = np.random.RandomState(3)
    H, W = 1080, 1920
    img = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    n = 300
    x0, y0 = rng.randint(0, W - 2, n), rng.randint(0, H - 2, n)
    boxes = np.stack([x0, y0, x0 + rng.randint(1, 200, n), y0 + rng.randint(1, 400, n)], axis=1).astype(np.int32)
    got = mot3d.color_histograms(img, boxes).cpu().numpy()
    for b in range(0, n, 7):
        xmin, ymin, xmax, ymax = boxes[b]
        ref = np.array(color_lab.color_histogram(img[ymin:ymax, xmin:xmax]))
        assert (got[b] == ref).all(), b
    # every box: counts add up to the clipped patch area (mean-subtracted histograms sum to ~0, exactly npix before)
    area = (np.minimum(boxes[:, 2], W) - boxes[:, 0]) * (np.minimum(boxes[:, 3], H) - boxes[:, 1])
    back = got[:, 0, :] + (area / 255.0)[:, None]
    assert np.allclose(back.sum(axis=1), area, rtol=0, atol=1e-6)
