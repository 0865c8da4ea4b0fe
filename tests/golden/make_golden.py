"""Regenerates tests/golden/fixture_frame.npz and tests/golden/golden_hashes.json.

Runs ONLY in the build container (needs /root/reference/benchmark/*.png and cv2 4.13).
The vectors are computed with REAL OpenCV primitives (cv2.GaussianBlur / Sobel / phase /
convertScaleAbs / pyrDown / medianBlur) plus plain numpy for the parts of
opencv_contrib rgbd/linemod.cpp that cv2-headless does not ship (hysteresis vote, spread,
response LUT, linearize, the normal-estimation index triple).  It does NOT import oracle/.
The resulting hashes are the ones SURVEY.md §8c lists as G1..G6 — the test-suite pins the
oracle against them (tests/test_oracle_golden.py).

Inputs are the reference's only shipped fixture: benchmark/img0.png, benchmark/depth0.png
(reference: detector.cpp:12,:25-26).
"""
import hashlib
import json
import os
import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/benchmark"
sha = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def sim_lut(circular):
    out = np.zeros(256, np.uint8)
    for ori in range(8):
        for half in range(2):
            for nib in range(16):
                best = 0
                for b in range(4):
                    if nib >> b & 1:
                        d = abs(ori - (b + 4 * half))
                        if circular:
                            d = min(d, 8 - d)
                        best = max(best, max(0, 4 - d))
                out[32 * ori + 16 * half + nib] = best
    return out


def cg_quantize(bgr, weak=10.0):
    sm = cv2.GaussianBlur(bgr, (7, 7), 0, 0, borderType=cv2.BORDER_REPLICATE)
    dx = cv2.Sobel(sm, cv2.CV_16S, 1, 0, ksize=3, borderType=cv2.BORDER_REPLICATE).astype(np.int32)
    dy = cv2.Sobel(sm, cv2.CV_16S, 0, 1, ksize=3, borderType=cv2.BORDER_REPLICATE).astype(np.int32)
    m = dx * dx + dy * dy
    pick = np.where((m[..., 0] >= m[..., 1]) & (m[..., 0] >= m[..., 2]), 0,
                    np.where((m[..., 1] >= m[..., 0]) & (m[..., 1] >= m[..., 2]), 1, 2))
    ii, jj = np.indices(pick.shape)
    gx = dx[ii, jj, pick].astype(np.float32); gy = dy[ii, jj, pick].astype(np.float32)
    mag = (gx * gx + gy * gy).astype(np.float32)
    ang = cv2.phase(gx, gy, angleInDegrees=True)
    q = cv2.convertScaleAbs(ang, alpha=16.0 / 360.0)
    q[0, :] = 0; q[-1, :] = 0; q[:, 0] = 0; q[:, -1] = 0
    q[1:-1, 1:-1] &= 7
    H, W = q.shape
    hist = np.zeros((8, H - 2, W - 2), np.int32)
    for di in range(3):
        for dj in range(3):
            patch = q[di:di + H - 2, dj:dj + W - 2]
            for b in range(8):
                hist[b] += patch == b
    idx = hist.argmax(0)            # first maximum, like `max_votes < histogram[i]`
    votes = hist.max(0)
    out = np.zeros_like(q)
    ok = (mag[1:-1, 1:-1] > np.float32(weak * weak)) & (votes >= 5)
    out[1:-1, 1:-1] = np.where(ok, (1 << idx).astype(np.uint8), 0)
    return out, mag


def spread(q, T):
    out = np.zeros_like(q)
    H, W = q.shape
    for r in range(T):
        for c in range(T):
            out[:H - r, :W - c] |= q[r:, c:]
    return out


def response(sp, lut):
    return np.stack([np.maximum(lut[32 * o + (sp & 15)], lut[32 * o + 16 + (sp >> 4)]) for o in range(8)])


def linearize(resp, T):
    return np.stack([resp[r0::T, c0::T].reshape(-1) for r0 in range(T) for c0 in range(T)])


def dn_indices(depth, dist_thr=2000, diff_thr=50):
    H, W = depth.shape
    d = depth.astype(np.int64)
    r = 5
    ys, xs = slice(r, H - r - 1), slice(r, W - r - 1)
    dc = d[ys, xs]
    A0 = np.zeros_like(dc); A1 = np.zeros_like(dc); A3 = np.zeros_like(dc); b0 = np.zeros_like(dc); b1 = np.zeros_like(dc)
    for (i, j) in [(-r, -r), (0, -r), (r, -r), (-r, 0), (r, 0), (-r, r), (0, r), (r, r)]:
        nb = d[r + j:H - r - 1 + j, r + i:W - r - 1 + i]
        delta = nb - dc
        f = (np.abs(delta) < diff_thr).astype(np.int64)
        A0 += f * i * i; A1 += f * i * j; A3 += f * j * j; b0 += f * i * delta; b1 += f * j * delta
    det = A0 * A3 - A1 * A1; ddx = A3 * b0 - A1 * b1; ddy = -A1 * b0 + A0 * b1
    nx = (1150 * ddx).astype(np.float32); ny = (1150 * ddy).astype(np.float32); nz = (-det * dc).astype(np.float32)
    s = np.sqrt(((nx * nx).astype(np.float32) + (ny * ny).astype(np.float32)).astype(np.float32) + (nz * nz).astype(np.float32)).astype(np.float32)
    valid = (dc < dist_thr) & (s > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (np.float32(1.0) / s).astype(np.float32)
        v1 = ((nx * inv).astype(np.float32) * np.float32(10) + np.float32(10)).astype(np.float32)
        v2 = ((ny * inv).astype(np.float32) * np.float32(10) + np.float32(10)).astype(np.float32)
        v3 = ((nz * inv).astype(np.float32) * np.float32(20) + np.float32(20)).astype(np.float32)
    out = np.full((3, H, W), -1, np.int8)
    for k, v in enumerate((v1, v2, v3)):
        vi = np.where(valid, np.nan_to_num(v).astype(np.int32), -1)   # C truncation toward zero
        out[k, ys, xs] = vi
    return out


def main():
    img = cv2.imread(os.path.join(REF, "img0.png"))
    dep = cv2.imread(os.path.join(REF, "depth0.png"), cv2.IMREAD_ANYDEPTH)
    np.savez_compressed(os.path.join(HERE, "fixture_frame.npz"), bgr=img, depth=dep)
    G = {"input": {"bgr_sha1": sha(img), "depth_sha1": sha(dep)}}
    # G1: exhaustive (dx,dy) -> label table through cv2.phase + convertScaleAbs
    dyy, dxx = np.meshgrid(np.arange(-1020, 1021, dtype=np.float32), np.arange(-1020, 1021, dtype=np.float32), indexing="ij")
    tab = (cv2.convertScaleAbs(cv2.phase(dxx, dyy, angleInDegrees=True), alpha=16.0 / 360.0) & 7).astype(np.uint8)
    G["G1_label_table"] = {"sha1": sha(tab), "hist": np.bincount(tab.ravel(), minlength=8).tolist()}
    for name, bgr, T in (("G2G3_L0_T5", img, 5), ("G4_L1_T8", cv2.pyrDown(img, dstsize=(img.shape[1] // 2, img.shape[0] // 2)), 8)):
        q, mag = cg_quantize(bgr)
        sp = spread(q, T)
        e = {"quantized_sha1": sha(q), "nnz": int((q > 0).sum()), "label_counts": [int((q == (1 << i)).sum()) for i in range(8)],
             "magnitude_sha1": sha(mag), "spread_sha1": sha(sp)}
        for circ in (0, 1):
            resp = response(sp, sim_lut(circ))
            lm = np.stack([linearize(resp[o], T) for o in range(8)])
            e["lut%d" % circ] = {"response_sha1": sha(resp), "linmem_sha1": sha(lm), "linmem_shape": list(lm.shape)}
        G[name] = e
    idx = dn_indices(dep)
    G["G5_dn_indices"] = {"sha1": sha(idx), "valid_fraction": float((idx[0, 5:474, 5:634] >= 0).mean()), "max": int(idx.max())}
    # DepthNormal L0 quantised map under the stand-in NORMAL_LUT is produced by tools/make_normal_lut.py (not an OpenCV pin).
    G["G6_similarity_lut"] = {"circular_sha1": sha(sim_lut(1)), "circular_sum": int(sim_lut(1).sum()),
                              "linear_sha1": sha(sim_lut(0)), "linear_sum": int(sim_lut(0).sum())}
    G["median5_check"] = "cv2.medianBlur compared directly in tests/test_oracle_primitives.py"
    json.dump(G, open(os.path.join(HERE, "golden_hashes.json"), "w"), indent=1)
    print(json.dumps(G, indent=1))


if __name__ == "__main__":
    main()
