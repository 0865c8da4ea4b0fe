"""Second, independent restatement of upstream's addTemplate side (ColorGradient /
DepthNormal extractTemplate, selectScatteredFeatures, cropTemplates) written from the appendix with REAL OpenCV
for erode / distanceTransform / resize(NEAREST), compared with the C++ oracle's add_template.  The quantized maps
and magnitudes it starts from are the oracle primitives that the golden hashes G2/G4/G5 pin.  CPU only."""
import math
import numpy as np
import pytest

from oracle import oracle as O
from line_mod_pipeline_b200 import synth

cv2 = pytest.importorskip("cv2")
F32 = np.float32


def select_scattered(cands, n, distance):
    """selectScatteredFeatures: cycle through the sorted candidates, keep one if it is at least `distance` away from
    every feature kept so far, lower the distance by 1.0f at every wrap-around.  cands: (x, y, label, score) sorted.
    Same greedy order as upstream's loop, evaluated with a running "squared distance to the nearest kept feature"
    per candidate so that whole fruitless passes collapse into one step."""
    xy = np.array([(c[0], c[1]) for c in cands], np.int64)
    near = np.full(len(cands), np.iinfo(np.int64).max // 4, np.int64)   # squared distance to the nearest kept feature
    out = []
    distance = F32(distance)
    dsq = F32(distance * distance)
    i = 0
    while len(out) < n:
        ok = np.nonzero(near[i:].astype(np.float32) >= dsq)[0]
        if len(ok):
            k = i + int(ok[0])
            out.append((cands[k][0], cands[k][1], cands[k][2]))
            d = xy - xy[k]
            near = np.minimum(near, d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
            i = k + 1
            if i == len(cands):
                i = 0
                distance = F32(distance - F32(1.0)); dsq = F32(distance * distance)
        else:                                    # nothing left in this pass: wrap around
            i = 0
            distance = F32(distance - F32(1.0)); dsq = F32(distance * distance)
    return out


def extract_cg(q, mag, mask, n, strong=55.0):
    if mask is not None:
        local = cv2.subtract(mask, cv2.erode(mask, None, iterations=1, borderType=cv2.BORDER_REPLICATE))
    else:
        local = np.full(q.shape, 255, np.uint8)
    ys, xs = np.nonzero((local != 0) & (q > 0) & (mag > F32(strong * strong)))       # raster order
    cands = [(int(x), int(y), int(math.log2(q[y, x])), F32(mag[y, x])) for y, x in zip(ys, xs)]
    if n <= 0 or len(cands) < n:
        return None
    cands.sort(key=lambda c: -float(c[3]))                                           # stable, score descending
    return select_scattered(cands, n, F32(len(cands) // n + 1))


def extract_dn(q, mask, n, extract_threshold):
    if mask is not None:
        local = cv2.erode(mask, None, iterations=2, borderType=cv2.BORDER_REPLICATE)
    else:
        local = np.full(q.shape, 255, np.uint8)
    dist = []
    for i in range(8):
        temp = np.where(local != 0, q & (1 << i), 0).astype(np.uint8)
        dist.append(cv2.distanceTransform(temp, cv2.DIST_C, 3))
    counts = [0] * 8
    cands = []
    ys, xs = np.nonzero((local != 0) & (q != 0) & (q != 255))
    for y, x in zip(ys, xs):
        label = int(math.log2(q[y, x]))
        score = dist[label][y, x]
        if score >= F32(extract_threshold):
            cands.append([int(x), int(y), label, F32(score)])
            counts[label] += 1
    if n <= 0 or len(cands) < n:
        return None
    cands = [(x, y, l, F32(s / F32(counts[l]))) for x, y, l, s in cands]
    cands.sort(key=lambda c: -float(c[3]))
    area = int(np.count_nonzero(local)) if mask is not None else q.size
    distance = F32(F32(math.sqrt(F32(area))) / F32(math.sqrt(F32(n))) + F32(1.5))
    return select_scattered(cands, n, distance)


def np_add_template(bgr, depth, mask, lut, L=2, nf=63):
    """-> None on failure, else (templates [l*2+m] = dict(width, height, pyramid_level, features), bb)."""
    tp = [None] * (2 * L)
    src, m, n = bgr, mask, nf
    for l in range(L):
        if l > 0:
            src = cv2.pyrDown(src)
            n //= 2
            if m is not None:
                m = cv2.resize(m, (m.shape[1] // 2, m.shape[0] // 2), interpolation=cv2.INTER_NEAREST)
        q, mag = O.cg_quantize(src)
        f = extract_cg(q, mag, m, n)
        if f is None:
            return None
        tp[l * 2] = dict(pyramid_level=l, features=f)
    q, m, n, thr = O.dn_quantize(depth, lut), mask, nf, 2
    for l in range(L):
        if l > 0:
            q = cv2.resize(q, (q.shape[1] // 2, q.shape[0] // 2), interpolation=cv2.INTER_NEAREST)
            n //= 2
            thr //= 2
            if m is not None:
                m = cv2.resize(m, (m.shape[1] // 2, m.shape[0] // 2), interpolation=cv2.INTER_NEAREST)
        f = extract_dn(q, m, n, thr)
        if f is None:
            return None
        tp[l * 2 + 1] = dict(pyramid_level=l, features=f)
    # cropTemplates
    xs = [x << t["pyramid_level"] for t in tp for x, _, _ in t["features"]]
    ys = [y << t["pyramid_level"] for t in tp for _, y, _ in t["features"]]
    min_x, max_x, min_y, max_y = min(xs), max(xs), min(ys), max(ys)
    if min_x % 2 == 1:
        min_x -= 1
    if min_y % 2 == 1:
        min_y -= 1
    for t in tp:
        lv = t["pyramid_level"]
        t["width"], t["height"] = (max_x - min_x) >> lv, (max_y - min_y) >> lv
        t["features"] = [(x - (min_x >> lv), y - (min_y >> lv), l) for x, y, l in t["features"]]
    return tp, (min_x, min_y, max_x - min_x, max_y - min_y)


@pytest.mark.parametrize("which", ["fixture", "synthetic"])
def test_numpy_addtemplate_equals_oracle(which, fixture_frame):
    lut = synth.default_normal_lut()
    bgr, depth = fixture_frame if which == "fixture" else synth.make_frame(2)
    masks = synth.planted_masks(10, seed=21) + (synth.object_masks(2)[:6] if which == "synthetic" else [])
    masks += [None, np.zeros((480, 640), np.uint8)]
    ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], [5, 8], normal_lut=lut)
    ok = fail = 0
    for i, m in enumerate(masks):
        want = np_add_template(bgr, depth, m, lut)
        tid, bb = ora.add_template([bgr, depth], "obj", m)
        if want is None:
            assert tid == -1, "mask %d: the oracle extracted a template, the restatement did not" % i
            fail += 1
            continue
        assert tid == ok, "mask %d: template id %d, expected %d" % (i, tid, ok)
        ok += 1
        tp, wbb = want
        assert tuple(bb) == wbb
        got = O.decode_pyramid(ora.get_template_flat("obj", tid))
        for a, b in zip(got, tp):
            assert (a["width"], a["height"], a["pyramid_level"]) == (b["width"], b["height"], b["pyramid_level"])
            assert [tuple(int(v) for v in f) for f in a["features"]] == b["features"]
    assert ok >= 5 and fail >= 1
