"""Oracle primitives vs REAL OpenCV (cv2 main modules) on random inputs — bit-exact.

These are the imgproc/core functions upstream linemod.cpp calls (SURVEY.md §8c row
"Primitive-level oracles available now").
"""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import oracle as O

RNG = np.random.default_rng(7)
SHAPES = [(48, 64), (97, 131), (33, 20)]


@pytest.mark.parametrize("shape", SHAPES)
def test_gauss7_sobel(shape):
    img = RNG.integers(0, 256, shape + (3,), dtype=np.uint8)
    g = O.gauss7(img)
    assert np.array_equal(g, cv2.GaussianBlur(img, (7, 7), 0, 0, borderType=cv2.BORDER_REPLICATE))
    dx, dy = O.sobel3(g)
    assert np.array_equal(dx, cv2.Sobel(g, cv2.CV_16S, 1, 0, ksize=3, borderType=cv2.BORDER_REPLICATE))
    assert np.array_equal(dy, cv2.Sobel(g, cv2.CV_16S, 0, 1, ksize=3, borderType=cv2.BORDER_REPLICATE))


@pytest.mark.parametrize("shape", SHAPES + [(1080, 1920)])
def test_pyrdown(shape):
    img = RNG.integers(0, 256, shape + (3,), dtype=np.uint8)
    assert np.array_equal(O.pyrdown(img), cv2.pyrDown(img, dstsize=(shape[1] // 2, shape[0] // 2)))


def test_fast_atan2_matches_cv2_phase():
    y = RNG.integers(-1020, 1021, 200000).astype(np.float32)
    x = RNG.integers(-1020, 1021, 200000).astype(np.float32)
    assert np.array_equal(O.fast_atan2(y, x, 1), cv2.phase(x, y, angleInDegrees=True).ravel())


def test_median_erode_dist_resize():
    m = RNG.choice(np.array([0, 1, 2, 4, 8, 16, 32, 64, 128], np.uint8), (77, 93))
    assert np.array_equal(O.median5(m), cv2.medianBlur(m, 5))
    mk = (RNG.random((60, 80)) > 0.3).astype(np.uint8) * 255
    assert np.array_equal(O.erode3(mk), cv2.erode(mk, None, borderType=cv2.BORDER_REPLICATE))
    assert np.array_equal(O.dist_c(mk), cv2.distanceTransform(mk, cv2.DIST_C, 3))
    blob = np.zeros((50, 70), np.uint8); blob[10:40, 15:60] = 4
    assert np.array_equal(O.dist_c(blob), cv2.distanceTransform(blob, cv2.DIST_C, 3))
    ones = np.full((20, 30), 255, np.uint8)
    assert np.array_equal(O.dist_c(ones), cv2.distanceTransform(ones, cv2.DIST_C, 3))
    mm = RNG.integers(0, 255, (61, 83), dtype=np.uint8)
    assert np.array_equal(O.resize_nn(mm, 30, 41), cv2.resize(mm, (41, 30), interpolation=cv2.INTER_NEAREST))
    assert np.array_equal(O.resize_nn(mm[:60, :82], 30, 41), mm[:60:2, :82:2])


def test_integer_label_rule():
    """The integer-compare orientation rule of csrc/kernels_color.cu (label_from_gradient) against the oracle's
    quantize(fastAtan2) & 7 table over ALL 2041^2 Sobel pairs (SURVEY golden G1 pins that table's sha1)."""
    import hashlib
    tab = np.empty((2041, 2041), np.uint8)
    O.lib().lmo_label_table(tab.ctypes.data_as(O.C.c_void_p), 1)
    assert hashlib.sha1(tab.tobytes()).hexdigest() == "521625b4366ecc3f18d4997fe506a3eb147af327"
    d = np.arange(-1020, 1021, dtype=np.int64)
    DY, DX = np.meshgrid(d, d, indexing="ij")
    ax, ay = np.abs(DX), np.abs(DY)
    mn, mx = np.minimum(ax, ay), np.maximum(ax, ay)
    N1, N2 = 208571, 700819
    assert int((mn << 20).max()) < 2 ** 31 and int((mx * N2).max()) < 2 ** 31      # the kernel's 32-bit products never overflow
    k = ((mn << 20) > mx * N1).astype(np.int64) + ((mn << 20) > mx * N2)
    t = np.where(ay > ax, 4 - k, k)
    lab = np.where((DX ^ DY) < 0, (8 - t) & 7, t)
    assert np.array_equal(lab.astype(np.uint8), tab)
