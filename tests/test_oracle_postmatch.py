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
