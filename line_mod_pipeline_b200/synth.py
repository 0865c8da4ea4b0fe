"""Synthetic RGB-D frames and template sets (numpy only) — the workload SURVEY.md §8d defines.

Frame (seed = 1234 + frame_idx): mid-grey base + 40 random filled rotated rectangles / ellipses
(uniform BGR, 20-200 px) + two low-frequency sinusoidal gratings (amplitude 20) + Gaussian noise
(sigma 2), clipped to u8.  Depth: background plane 1200 mm; each shape a tilted plane 500-1100 mm
(slope <= 0.5 mm/px); 5 % of 8x8 blocks are 0-holes; every non-zero depth >= 400 mm.
Random templates (seed 99): bbox w,h even in [60,200]; per modality nf features uniform in
[0,w]x[0,h] with one on each bbox edge, labels uniform 0..7; level l: (w>>l, h>>l), nf>>l features.
"""
import numpy as np


def make_frame(idx=0, rows=480, cols=640, n_shapes=40, return_owner=False):
    rng = np.random.default_rng(1234 + idx)
    yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float32)
    color = np.full((rows, cols, 3), 128.0, np.float32)
    depth = np.full((rows, cols), 1200.0, np.float32)
    owner = np.full((rows, cols), -1, np.int32)      # index of the topmost shape per pixel
    for k in range(n_shapes):
        cx, cy = rng.uniform(0, cols), rng.uniform(0, rows)
        a, b = rng.uniform(10, 100), rng.uniform(10, 100)
        th = rng.uniform(0, np.pi)
        u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        if rng.random() < 0.5:
            m = (u / a) ** 2 + (v / b) ** 2 <= 1.0
        else:
            m = (np.abs(u) <= a) & (np.abs(v) <= b)
        col = rng.uniform(20, 235, 3)
        color[m] = col
        owner[m] = k
        d0 = rng.uniform(500, 1100)
        sx, sy = rng.uniform(-0.5, 0.5, 2)
        depth[m] = (d0 + sx * (xx - cx) + sy * (yy - cy))[m]
    for _ in range(2):
        fx, fy = rng.uniform(0.005, 0.02, 2)
        ph = rng.uniform(0, 2 * np.pi)
        color += (20.0 * np.sin(2 * np.pi * (fx * xx + fy * yy) + ph))[..., None]
    color += rng.normal(0, 2.0, color.shape)
    bgr = np.clip(np.rint(color), 0, 255).astype(np.uint8)
    depth = np.clip(depth, 400, 1999)
    holes = rng.random(((rows + 7) // 8, (cols + 7) // 8)) < 0.05
    depth[np.kron(holes, np.ones((8, 8), bool))[:rows, :cols]] = 0
    if return_owner:
        return bgr, np.rint(depth).astype(np.uint16), owner
    return bgr, np.rint(depth).astype(np.uint16)


def object_masks(idx=0, rows=480, cols=640, n_shapes=40, min_px=1500, margin=12):
    """Object masks for planting templates: the visible region of every shape of frame `idx` that is large
    enough and away from the image border (uint8 0/255), i.e. silhouettes that coincide with real edges."""
    _, _, owner = make_frame(idx, rows, cols, n_shapes, return_owner=True)
    out = []
    for k in range(n_shapes):
        m = owner == k
        if m.sum() < min_px:
            continue
        ys, xs = np.nonzero(m)
        if ys.min() < margin or xs.min() < margin or ys.max() >= rows - margin or xs.max() >= cols - margin:
            continue
        out.append(m.astype(np.uint8) * 255)
    return out


def random_template_pyramid(rng, n_modalities, levels, nf0=63, wh_range=(60, 200)):
    """-> list (level*M + modality) of dict(width,height,pyramid_level,features)."""
    w = int(rng.integers(wh_range[0] // 2, wh_range[1] // 2 + 1)) * 2
    h = int(rng.integers(wh_range[0] // 2, wh_range[1] // 2 + 1)) * 2
    out = []
    for l in range(levels):
        wl, hl, nf = w >> l, h >> l, max(4, nf0 >> l)
        for m in range(n_modalities):
            x = rng.integers(0, wl + 1, nf)
            y = rng.integers(0, hl + 1, nf)
            x[0], x[1] = 0, wl
            y[2], y[3] = 0, hl
            lab = rng.integers(0, 8, nf)
            out.append(dict(width=wl, height=hl, pyramid_level=l, features=np.stack([x, y, lab], 1).astype(np.int32)))
    return out


def random_templates(n, n_modalities=2, levels=2, seed=99, nf0=63, wh_range=(60, 200)):
    rng = np.random.default_rng(seed)
    return [random_template_pyramid(rng, n_modalities, levels, nf0, wh_range) for _ in range(n)]


def planted_masks(n, rows=480, cols=640, seed=7, size_range=(60, 200)):
    """Random rectangular / elliptic object masks (uint8 0/255) used to plant templates cut from a frame."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:rows, 0:cols]
    out = []
    for _ in range(n):
        w, h = rng.integers(size_range[0], size_range[1] + 1, 2)
        x0 = int(rng.integers(16, max(17, cols - w - 16)))
        y0 = int(rng.integers(16, max(17, rows - h - 16)))
        if rng.random() < 0.5:
            m = (xx >= x0) & (xx < x0 + w) & (yy >= y0) & (yy < y0 + h)
        else:
            m = ((xx - (x0 + w / 2)) / (w / 2)) ** 2 + ((yy - (y0 + h / 2)) / (h / 2)) ** 2 <= 1.0
        out.append(m.astype(np.uint8) * 255)
    return out


def default_normal_lut():
    """The stand-in NORMAL_LUT[20][20][20] (same integer rule as csrc/capi.cpp default_normal_lut)."""
    out = np.zeros((20, 20, 20), np.uint8)
    for v2 in range(20):
        for v1 in range(20):
            x, y = 2 * v1 - 19, 2 * v2 - 19
            ax, ay = abs(x), abs(y)
            s = (ax + ay) ** 2
            if s < 2 * ax * ax:
                b = 0 if x > 0 else 4
            elif s < 2 * ay * ay:
                b = 2 if y > 0 else 6
            else:
                b = (1 if y > 0 else 7) if x > 0 else (3 if y > 0 else 5)
            out[:, v2, v1] = 1 << b
    return out.reshape(-1)
