"""Generates tests/golden/lagergehaeuse_templates.yml.gz — the BASELINE config-1 template set.

Runs ONLY in the build container (needs /root/reference/models/lagergehaeuse.ply and cv2).
Restates the reference's template generation for its one shipped model (SURVEY.md §8c "Config-1 templates"):
  * viewpoints: CameraViewPoints::createVerticesForRotSym + removeSuperfluousVertices
    (src/CameraViewPoints.cpp:34-52,:75-82) -> 13 viewpoints x radii 500..1200 step 50
    (linemod_settings.yml:24-26, TemplateGenerator.cpp:47) = 195 renders;
  * render: pinhole fy=1045.69 for both axes, 640x480, lookAt(pos, 0, +Y), depth in mm (u16), binary colour
    (OpenglRender.cpp:3-12,:334-345, shader/depth.fs) by a small numpy z-buffer rasteriser standing in for OpenGL;
  * HighLevelLineMOD::addTemplate (src/HighLevelLinemod.cpp:68-110): threshold, 10 in-plane rotations
    (-45..45 step 10, warpAffine), erode, then Detector::addTemplate — here the ORACLE's addTemplate.
Pixel-exact equality with the author's GL output is neither possible nor required: config-1 parity is
GPU-vs-oracle on this same template set.  The file is written with the product's persistence code
(host-only) in the reference's linemod_templates.yml.gz layout.
"""
import math
import os
import sys
import time
import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O
import line_mod_pipeline_b200 as lm
from line_mod_pipeline_b200 import synth

W, H, FY = 640, 480, 1045.69141
NEAR, FAR = 100.0, 10000.0


def load_ply(path):
    with open(path) as f:
        assert f.readline().strip() == "ply"
        nv = nf = 0
        while True:
            l = f.readline().strip()
            if l.startswith("element vertex"):
                nv = int(l.split()[-1])
            elif l.startswith("element face"):
                nf = int(l.split()[-1])
            elif l == "end_header":
                break
        v = np.array([[float(x) for x in f.readline().split()[:3]] for _ in range(nv)], np.float64)
        faces = []
        for _ in range(nf):
            p = [int(x) for x in f.readline().split()]
            for k in range(1, p[0] - 1):      # triangulate fans like assimp
                faces.append((p[1], p[1 + k], p[2 + k]))
    return v, np.array(faces, np.int64)


def look_at(eye):
    eye = np.array(eye, np.float64)
    if eye[0] == 0 and eye[2] == 0:
        eye[0] = eye[2] = 1e-6
    f = -eye / np.linalg.norm(eye)
    up = np.array([0.0, 1.0, 0.0])
    s = np.cross(f, up); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    R = np.stack([s, u, -f])
    return R, -R @ eye


def render(verts, faces, eye):
    """-> (depth u16 mm, colour u8 3ch) as the reference's GL passes would deliver them after the vertical flip."""
    R, t = look_at(eye)
    vc = verts @ R.T + t                     # camera space, looking down -z
    z = -vc[:, 2]
    u = W / 2 + FY * vc[:, 0] / z
    v = H / 2 - FY * vc[:, 1] / z
    depth = np.full((H, W), np.inf)
    for (a, b, c) in faces:
        if z[a] < NEAR or z[b] < NEAR or z[c] < NEAR:
            continue
        x0, x1 = int(math.floor(min(u[a], u[b], u[c]))), int(math.ceil(max(u[a], u[b], u[c])))
        y0, y1 = int(math.floor(min(v[a], v[b], v[c]))), int(math.ceil(max(v[a], v[b], v[c])))
        x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, W - 1), min(y1, H - 1)
        if x0 > x1 or y0 > y1:
            continue
        xs, ys = np.meshgrid(np.arange(x0, x1 + 1) + 0.5, np.arange(y0, y1 + 1) + 0.5)
        d = (v[b] - v[c]) * (u[a] - u[c]) + (u[c] - u[b]) * (v[a] - v[c])
        if abs(d) < 1e-12:
            continue
        l0 = ((v[b] - v[c]) * (xs - u[c]) + (u[c] - u[b]) * (ys - v[c])) / d
        l1 = ((v[c] - v[a]) * (xs - u[c]) + (u[a] - u[c]) * (ys - v[c])) / d
        l2 = 1 - l0 - l1
        inside = (l0 >= 0) & (l1 >= 0) & (l2 >= 0)
        if not inside.any():
            continue
        zi = 1.0 / (l0 / z[a] + l1 / z[b] + l2 / z[c])   # perspective-correct depth
        sub = depth[y0:y1 + 1, x0:x1 + 1]
        upd = inside & (zi < sub)
        sub[upd] = zi[upd]
    hit = np.isfinite(depth)
    d16 = np.zeros((H, W), np.uint16)
    d16[hit] = np.clip(np.rint(depth[hit]), 0, 65535).astype(np.uint16)   # depth.fs: linear z / far / 6.5535 of a u16 target = mm
    col = np.zeros((H, W, 3), np.uint8)
    col[hit] = 255                                                          # PLY without colours -> white (ModelImporter.cpp:53-71)
    return d16, col


def viewpoints(radius, subdivisions=3):
    out = []
    i = 0
    step = 60 / 2 ** subdivisions               # 7.5, added to a uint16_t: truncates every iteration
    while i < 360:
        vtx = (0.0, math.sin(i * math.pi / 180.0) * radius, math.cos(i * math.pi / 180.0) * radius)
        if vtx[0] >= 0 and vtx[1] >= 0 and vtx[2] >= 0:      # planes of symmetry (1,1,1)
            out.append(vtx)
        i = int(i + step)
    return out


def main():
    verts, faces = load_ply("/root/reference/models/lagergehaeuse.ply")
    print("mesh", verts.shape, faces.shape, verts.min(0), verts.max(0))
    lut = synth.default_normal_lut()
    ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], [5, 8], normal_lut=lut)
    rots = [cv2.getRotationMatrix2D((W // 2, H // 2), a, 1.0) for a in range(-45, 46, 10)]
    t0 = time.time()
    n_ok = n_fail = 0
    for radius in range(500, 1201, 50):
        for eye in viewpoints(float(radius)):
            depth, colour = render(verts, faces, eye)
            _, binary = cv2.threshold(colour, 1, 255, cv2.THRESH_BINARY)
            _, mask = cv2.threshold(depth, 1, 65535, cv2.THRESH_BINARY)
            mask = cv2.convertScaleAbs(mask)      # convertTo(CV_8UC1) saturates to 255
            for Rm in rots:
                m_r = cv2.warpAffine(mask, Rm, (W, H))
                c_r = cv2.warpAffine(binary, Rm, (W, H))
                d_r = cv2.warpAffine(depth, Rm, (W, H))
                m_r = cv2.erode(m_r, None)
                tid, bb = ora.add_template([c_r, d_r], "lagergehaeuse.ply", m_r)
                if tid < 0:
                    n_fail += 1
                    break                         # reference: addTemplate returns false -> viewpoint aborted (HighLevelLinemod.cpp:97-101)
                n_ok += 1
        print("radius", radius, "templates", n_ok, "failed viewpoints", n_fail, "%.0fs" % (time.time() - t0), flush=True)
    det = lm.getDefaultLINEMOD()
    for t in range(ora.num_templates("lagergehaeuse.ply")):
        det.addSyntheticTemplate(O.decode_pyramid(ora.get_template_flat("lagergehaeuse.ply", t)), "lagergehaeuse.ply")
    out = os.path.join(HERE, "lagergehaeuse_templates.yml.gz")
    det.write(out)
    print("wrote", out, det.numTemplates(), "templates,", os.path.getsize(out), "bytes")
    z = np.load(os.path.join(HERE, "fixture_frame.npz"))
    for thr in (80.0, 70.0, 60.0):
        m = ora.match([z["bgr"], z["depth"]], thr, threads=8).matches(0)
        print("fixture frame, threshold", thr, ":", len(m), "matches; best:", m[:3].tolist() if len(m) else None)


if __name__ == "__main__":
    main()
