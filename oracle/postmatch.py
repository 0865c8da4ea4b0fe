"""CPU restatement (numpy, test infrastructure) of the image operations behind HighLevelLineMOD's post-match colour
check — SURVEY.md §8f-3, the next row after the match path:

  detectTemplate   cvtColor(BGR2HSV) + inRange(lower, upper)              src/HighLevelLinemod.cpp:159-161
  templateMask     convexHull(features + match offset) + fillPoly(255)    src/HighLevelLinemod.cpp:113-135
  colorCheck       countNonZero(hue & mask) * 100 / countNonZero(mask) > percentToPassCheck   :424-434

The OpenCV primitives are restated from their published algorithms and pinned bit-for-bit against cv2 4.13 by
tests/test_oracle_postmatch.py: 8-bit BGR2HSV (fixed-point tables, hsv_shift 12), convexHull (vertex set; collinear
points dropped), fillPoly (8-connected boundary lines, left to right, + the fixed-point scanline fill).  Polygons are
assumed to lie inside the image, which holds for match positions the detector returns (cv::Line clips a line before
rasterising it, which changes its pixels).  No product code uses this module; the CUDA side of the row is not built yet.
"""
import numpy as np

_HSV_SHIFT = 12
_SDIV = np.zeros(256, np.int64)
_HDIV180 = np.zeros(256, np.int64)
for _i in range(1, 256):
    _SDIV[_i] = int(round((255 << _HSV_SHIFT) / (1.0 * _i)))
    _HDIV180[_i] = int(round((180 << _HSV_SHIFT) / (6.0 * _i)))
_XY_SHIFT = 16


def bgr2hsv(bgr):
    """cv::cvtColor(COLOR_BGR2HSV) for 8-bit images (H in 0..179)."""
    b = bgr[..., 0].astype(np.int64); g = bgr[..., 1].astype(np.int64); r = bgr[..., 2].astype(np.int64)
    v = np.maximum(np.maximum(b, g), r)
    diff = v - np.minimum(np.minimum(b, g), r)
    vr = np.where(v == r, -1, 0); vg = np.where(v == g, -1, 0)
    s = (diff * _SDIV[v] + (1 << (_HSV_SHIFT - 1))) >> _HSV_SHIFT
    h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))))
    h = (h * _HDIV180[diff] + (1 << (_HSV_SHIFT - 1))) >> _HSV_SHIFT
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


def in_range(img, lower, upper):
    """cv::inRange on a 3-channel 8-bit image -> 0/255 mask."""
    lo = np.asarray(lower).reshape(1, 1, 3); hi = np.asarray(upper).reshape(1, 1, 3)
    return (np.all((img >= lo) & (img <= hi), -1) * 255).astype(np.uint8)


def convex_hull(points):
    """Vertices of the convex hull of integer points (collinear points dropped), counter-clockwise from the
    lexicographically smallest: the vertex set cv::convexHull returns (its order differs; fillPoly does not care)."""
    P = sorted(set((int(x), int(y)) for x, y in points))
    if len(P) <= 2:
        return P

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])
    lo, up = [], []
    for p in P:
        while len(lo) >= 2 and cross(lo[-2], lo[-1], p) <= 0:
            lo.pop()
        lo.append(p)
    for p in reversed(P):
        while len(up) >= 2 and cross(up[-2], up[-1], p) <= 0:
            up.pop()
        up.append(p)
    return lo[:-1] + up[:-1]


def _line8(img, p1, p2, val):
    """cv::Line, connectivity 8: LineIterator with leftToRight = true (end points inside the image)."""
    (x1, y1), (x2, y2) = p1, p2
    dx, dy = x2 - x1, y2 - y1
    sy = 1
    if dx < 0:
        dx, dy = -dx, -dy
        x1, y1 = x2, y2
    if dy < 0:
        dy, sy = -dy, -1
    vert = dy > dx
    if vert:
        dx, dy = dy, dx
    err, plus, minus = dx - 2 * dy, 2 * dx, -2 * dy
    x, y = x1, y1
    for _ in range(dx + 1):
        img[y, x] = val
        neg = err < 0
        err += minus + (plus if neg else 0)
        if vert:
            y += sy
            x += 1 if neg else 0
        else:
            x += 1
            y += sy if neg else 0


def fill_poly(img, pts, val=255):
    """cv::fillPoly(img, [pts], val) (line type 8, shift 0) for a polygon inside the image: CollectPolyEdges draws the
    boundary, FillEdgeCollection fills scanline spans between pairs of active edges with 16.16 fixed-point x."""
    n = len(pts)
    if n == 0:
        return
    edges = []
    p0 = pts[-1]
    for p1 in pts:
        _line8(img, p0, p1, val)
        x0, y0, x1, y1 = p0[0] << _XY_SHIFT, p0[1], p1[0] << _XY_SHIFT, p1[1]
        if y0 != y1:
            e = dict(y0=y0, y1=y1, x=x0) if y0 < y1 else dict(y0=y1, y1=y0, x=x1)
            num, den = x1 - x0, y1 - y0
            q = abs(num) // abs(den)                       # C++ integer division truncates toward zero
            e["dx"] = q if (num >= 0) == (den >= 0) else -q
            edges.append(e)
        p0 = p1
    if len(edges) < 2:
        return
    edges.sort(key=lambda e: (e["y0"], e["x"], e["dx"]))
    y_max = min(max(e["y1"] for e in edges), img.shape[0])
    active, i, y = [], 0, edges[0]["y0"]
    while y < y_max:
        active = [e for e in active if e["y1"] != y]
        while i < len(edges) and edges[i]["y0"] == y:
            active.append(edges[i]); i += 1
        active.sort(key=lambda e: e["x"])
        for k in range(0, len(active) - 1, 2):
            a, b = active[k], active[k + 1]
            xa, xb = min(a["x"], b["x"]), max(a["x"], b["x"])
            x1, x2 = (xa + (1 << _XY_SHIFT) - 1) >> _XY_SHIFT, xb >> _XY_SHIFT
            if y >= 0 and x1 < img.shape[1] and x2 >= 0:
                img[y, max(x1, 0):min(x2, img.shape[1] - 1) + 1] = val
            a["x"] += a["dx"]; b["x"] += b["dx"]
        y += 1


def template_mask(templates, num_modalities, match_x, match_y, rows, cols):
    """HighLevelLineMOD::templateMask: hull of the level-0 features of every modality, shifted to the match."""
    pts = [(int(x) + match_x, int(y) + match_y) for m in range(num_modalities) for x, y, _ in templates[m]["features"]]
    mask = np.zeros((rows, cols), np.uint8)
    fill_poly(mask, convex_hull(pts))
    return mask


def color_check(hue_mask, mask, percent_to_pass):
    """HighLevelLineMOD::colorCheck: integer percentage of mask pixels inside the colour range."""
    inside = int(np.count_nonzero(hue_mask & mask))
    return np.float32(inside * 100 // int(np.count_nonzero(mask))) > np.float32(percent_to_pass)
