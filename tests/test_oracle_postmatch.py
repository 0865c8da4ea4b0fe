"""Pins the restatement of the post-match colour check (oracle/postmatch.py, SURVEY.md §8f-3 groundwork) against real
OpenCV: BGR2HSV on a dense sample of the colour cube, convexHull + fillPoly on random point sets (incl. degenerate
ones), and the whole templateMask / colorCheck composition as the reference writes it (src/HighLevelLinemod.cpp:113-135,
:159-161, :424-434) on a synthetic frame with templates of the oracle.  CPU only."""
import random
import numpy as np
import pytest

from oracle import oracle as O
from oracle import postmatch as PM
from line_mod_pipeline_b200 import synth

cv2 = pytest.importorskip("cv2")


def test_bgr2hsv_equals_opencv():
    g, b = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    for r in list(range(0, 256, 5)) + [1, 2, 254, 255]:
        img = np.stack([b, g, np.full_like(b, r)], -1).astype(np.uint8)
        assert np.array_equal(PM.bgr2hsv(img), cv2.cvtColor(img, cv2.COLOR_BGR2HSV)), "red = %d" % r
    bgr, _ = synth.make_frame(3)
    hsv = cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV)
    lo, hi = (20, 30, 40), (120, 255, 200)
    assert np.array_equal(PM.in_range(PM.bgr2hsv(bgr), lo, hi), cv2.inRange(hsv, lo, hi))


def test_hull_and_fill_equal_opencv():
    rng = random.Random(7)
    for t in range(1500):
        k = rng.randrange(1, 60)
        span = rng.choice([3, 8, 79])
        off = 0 if span == 79 else 20
        pts = [(rng.randrange(0, span + 1) + off, rng.randrange(0, span + 1) + off) for _ in range(k)]
        ref_hull = cv2.convexHull(np.array(pts, np.int32))[:, 0, :]
        mine = PM.convex_hull(pts)
        assert set(map(tuple, ref_hull.tolist())) == set(mine)
        ref = np.zeros((80, 96), np.uint8)
        cv2.fillPoly(ref, [ref_hull], 255)
        got = np.zeros((80, 96), np.uint8)
        PM.fill_poly(got, mine)
        assert np.array_equal(ref, got), "point set %d" % t


def test_color_check_composition_equals_the_reference_formulation():
    ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], [5, 8], normal_lut=synth.default_normal_lut())
    bgr, depth = synth.make_frame(4)
    for m in synth.object_masks(4)[:10]:
        ora.add_template([bgr, depth], "obj", m)
    res = ora.match([bgr, depth], 85.0).matches(0)
    assert len(res) >= 5
    lower, upper = (0, 0, 90), (180, 255, 255)
    hue_ref = cv2.inRange(cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV), lower, upper)
    hue = PM.in_range(PM.bgr2hsv(bgr), lower, upper)
    assert np.array_equal(hue, hue_ref)
    verdicts = set()
    for i in range(min(len(res), 25)):
        tp = O.decode_pyramid(ora.get_template_flat("obj", int(res.template_id[i])))
        x, y = int(res.x[i]), int(res.y[i])
        # the reference: points = features of templates[0..M) + offset; convexHull; fillPoly; bitwise_and; ratio
        pts = np.array([(fx + x, fy + y) for m in range(2) for fx, fy, _ in tp[m]["features"]], np.int32)
        ref_mask = np.zeros((480, 640), np.uint8)
        cv2.fillPoly(ref_mask, [cv2.convexHull(pts)[:, 0, :]], 255)
        mask = PM.template_mask(tp, 2, x, y, 480, 640)
        assert np.array_equal(mask, ref_mask), "match %d" % i
        ratio = cv2.countNonZero(cv2.bitwise_and(hue_ref, ref_mask)) * 100 // cv2.countNonZero(ref_mask)
        for percent in (ratio - 0.5, float(ratio), ratio + 0.5):     # strict '>' on the integer percentage
            ref = np.float32(ratio) > np.float32(percent)
            assert bool(PM.color_check(hue, mask, percent)) == bool(ref)
            verdicts.add(bool(ref))
    assert verdicts == {True, False}


def test_median_depth_of_depth_check(tmp_path):
    """lmb200_postmatch_median_depth = HighLevelLineMOD::medianMat (src/HighLevelLinemod.cpp:336-349): the sequence it
    selects from is pinned against cv2 (threshold / subtract / saturating add / ROI), the (n/4)-th order statistic against
    numpy, and the reference's vec[n/5] — an element std::nth_element leaves in unspecified order — against the same
    libstdc++ call compiled separately."""
    import ctypes as C
    import subprocess
    import cv2
    import line_mod_pipeline_b200 as lm
    L = lm.capi.lib()
    rng = np.random.default_rng(11)
    depth = rng.integers(0, 1500, (120, 160)).astype(np.uint16)
    depth[rng.random(depth.shape) < 0.2] = 0
    depth[rng.random(depth.shape) < 0.02] = 1
    src = tmp_path / "nth.cpp"
    src.write_text("#include <algorithm>\n#include <cstdio>\n#include <vector>\n#include <cstdint>\n"
                   "int main(int c, char** v) { std::vector<uint16_t> a; unsigned x; FILE* f = fopen(v[1], \"r\");"
                   " while (fscanf(f, \"%u\", &x) == 1) a.push_back((uint16_t)x); fclose(f);"
                   " std::nth_element(a.begin(), a.begin() + a.size() / 4, a.end()); printf(\"%u\\n\", (unsigned)a[a.size() / 5]); }\n")
    exe = tmp_path / "nth"
    subprocess.run(["g++", "-O2", "-o", str(exe), str(src)], check=True)
    for bb in [(10, 20, 50, 40), (0, 0, 160, 120), (100, 90, 60, 30), (5, 5, 1, 1), (7, 3, 9, 2)]:
        x, y, w, h = bb
        _, inv = cv2.threshold(depth, 1, 65535, cv2.THRESH_BINARY)
        seq = cv2.add(depth, 65535 - inv)[y:y + h, x:x + w].reshape(-1)
        bb4 = (C.c_int * 4)(*bb)
        out = C.c_uint16(0)
        assert L.lmb200_postmatch_median_depth(depth.ctypes.data, 120, 160, 0, bb4, 4, C.byref(out)) == 0
        assert out.value == int(np.sort(seq)[len(seq) // 4])
        assert L.lmb200_postmatch_median_depth(depth.ctypes.data, 120, 160, 0, bb4, 5, C.byref(out)) == 0
        (tmp_path / "seq.txt").write_text(" ".join(str(int(v)) for v in seq))
        want = int(subprocess.run([str(exe), str(tmp_path / "seq.txt")], capture_output=True, text=True, check=True).stdout)
        assert out.value == want and out.value <= int(np.sort(seq)[len(seq) // 4])
    bad = (C.c_int * 4)(100, 100, 80, 40)                  # leaves the image: cv::Mat::operator()(Rect) asserts
    assert L.lmb200_postmatch_median_depth(depth.ctypes.data, 120, 160, 0, bad, 5, C.byref(out)) == lm.capi.E_INVALID
