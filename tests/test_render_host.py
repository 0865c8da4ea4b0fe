"""Headless rasteriser (lmb200_render_lookat / _pose / lmb200_load_ply) against the numpy z-buffer
rasteriser that stands in for the reference's OpenGL passes (tests/golden/make_config1_templates.py: it produced the
committed config-1 template set).  Both are float64; numpy's matrix products may fuse multiply-adds, so the tolerance
is: depth within 1 mm on every pixel both renderers hit, silhouettes differing on at most 1e-4 of the pixels.  CPU only."""
import importlib.util
import os
import time
import numpy as np
import pytest

from line_mod_pipeline_b200 import render as R, capi as K

HERE = os.path.dirname(os.path.abspath(__file__))


def _harness():
    pytest.importorskip("cv2")
    spec = importlib.util.spec_from_file_location("make_config1_templates", os.path.join(HERE, "golden", "make_config1_templates.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def torus_mesh(R0=60.0, r0=22.0, nu=36, nv=18):
    """A torus around the y axis with a box glued on: concave silhouettes, self-occlusion, quads and triangles."""
    u = np.linspace(0, 2 * np.pi, nu, endpoint=False); v = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    verts = np.stack([(R0 + r0 * np.cos(vv)) * np.cos(uu), r0 * np.sin(vv), (R0 + r0 * np.cos(vv)) * np.sin(uu)], -1).reshape(-1, 3)
    faces = []
    for i in range(nu):
        for j in range(nv):
            a, b = i * nv + j, i * nv + (j + 1) % nv
            c, d = ((i + 1) % nu) * nv + (j + 1) % nv, ((i + 1) % nu) * nv + j
            faces.append([a, b, c, d])
    box = np.array([[x, y, z] for x in (-15.0, 15.0) for y in (-40.0, 40.0) for z in (-15.0, 15.0)])
    o = len(verts)
    quads = [[0, 1, 3, 2], [4, 6, 7, 5], [0, 4, 5, 1], [2, 3, 7, 6], [0, 2, 6, 4], [1, 5, 7, 3]]
    faces += [[o + k for k in q] for q in quads]
    return np.concatenate([verts, box]), faces


def write_ply(path, verts, faces):
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment test mesh\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % (len(verts), len(faces)))
        for v in verts:
            f.write("%.9g %.9g %.9g\n" % tuple(v))
        for q in faces:
            f.write("%d %s\n" % (len(q), " ".join(str(int(k)) for k in q)))


def _compare(d, c, hd, hc, what):
    hit, hhit = d > 0, hd > 0
    assert np.count_nonzero(hit != hhit) <= 1e-4 * hit.size, "%s: silhouettes differ on %d pixels" % (what, np.count_nonzero(hit != hhit))
    both = hit & hhit
    assert both.sum() > 500, "%s: the model is not in view" % what
    assert np.abs(d[both].astype(np.int32) - hd[both].astype(np.int32)).max() <= 1, "%s: depth differs by more than 1 mm" % what
    assert np.array_equal(c[..., 0] > 0, hit) and np.array_equal(c[..., 0], c[..., 2]) and set(np.unique(c)) <= {0, 255}
    assert np.array_equal(hc[..., 0] > 0, hhit)


def test_ply_loader_and_views_match_the_numpy_rasteriser(tmp_path):
    h = _harness()
    verts, faces = torus_mesh()
    path = str(tmp_path / "torus.ply")
    write_ply(path, verts, faces)
    v, t = R.load_ply(path)
    hv, ht = h.load_ply(path)
    assert np.array_equal(v, hv) and np.array_equal(t, ht) and len(t) == 2 * len(faces)      # quads -> fans of two
    eyes = [(0.0, 0.0, 600.0), (0.0, 420.0, 420.0), (0.0, 700.0, 0.0), (300.0, 200.0, -500.0), (-350.0, -150.0, 260.0)]
    d, c = R.render_lookat(v, t, eyes)
    assert d.shape == (len(eyes), 480, 640) and c.shape == (len(eyes), 480, 640, 3)
    for i, eye in enumerate(eyes):
        hd, hc = h.render(hv, ht, eye)
        _compare(d[i], c[i], hd, hc, "eye %s" % (eye,))
    # straight down the y axis exercises the degenerate-up fix; the centre pixel sees the top of the box at 700 - 40 mm
    assert abs(int(d[2][240, 320]) - 660) <= 1
    # explicit model-view transforms give the same images as lookAt
    Rt = [h.look_at(e) for e in eyes]
    d2, c2 = R.render_pose(v, t, [r for r, _ in Rt], [tt for _, tt in Rt])
    assert np.abs(d2.astype(np.int32) - d.astype(np.int32)).max() <= 1 and np.count_nonzero((d2 > 0) != (d > 0)) <= 1e-4 * d.size
    # thread count does not change anything
    d1, c1 = R.render_lookat(v, t, eyes, threads=1)
    assert np.array_equal(d1, d) and np.array_equal(c1, c)


def test_reference_model_views(tmp_path):
    """The reference's own model at config-1 viewpoints (only where /root/reference is mounted: the build container)."""
    ply = "/root/reference/models/lagergehaeuse.ply"
    if not os.path.exists(ply):
        pytest.skip("reference checkout not present")
    h = _harness()
    v, t = R.load_ply(ply)
    hv, ht = h.load_ply(ply)
    assert np.array_equal(v, hv) and np.array_equal(t, ht)
    eyes = h.viewpoints(600.0)[::4] + h.viewpoints(1100.0)[1::5]
    t0 = time.perf_counter()
    d, c = R.render_lookat(v, t, eyes)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = [h.render(hv, ht, e) for e in eyes]
    dt_np = time.perf_counter() - t0
    for i, (hd, hc) in enumerate(ref):
        _compare(d[i], c[i], hd, hc, "viewpoint %d" % i)
    print("rendered %d views: %.1f ms/view (numpy harness %.1f ms/view)" % (len(eyes), 1e3 * dt / len(eyes), 1e3 * dt_np / len(eyes)))


def test_argument_errors(tmp_path):
    verts, faces = torus_mesh()
    path = str(tmp_path / "t.ply")
    write_ply(path, verts, faces)
    v, t = R.load_ply(path)
    with pytest.raises(R.RenderError) as e:
        R.load_ply(str(tmp_path / "missing.ply"))
    assert e.value.code == K.E_IO
    bad = t.copy(); bad[0, 0] = len(v)
    with pytest.raises(R.RenderError) as e:
        R.render_lookat(v, bad, [(0, 0, 500.0)])
    assert e.value.code == K.E_INVALID
    with pytest.raises(R.RenderError):
        R.render_lookat(v, t, [(0, 0, 500.0)], depth=False, colour=False)
    d, c = R.render_lookat(v, t, [(0, 0, 500.0)], colour=False)
    assert c is None and d.max() > 0
    # a model entirely behind the near plane renders nothing
    d, _ = R.render_lookat(v, t, [(0.0, 0.0, 90.0)], camera=R.Camera(near_mm=100.0), colour=False)
    assert d.max() == 0 or d[d > 0].min() >= 100


def _hodan_cv2(inp, gt, est, vis_thr=15, err_thr=20):
    """Benchmark::calculateErrorHodan + calculateVisibilityMasks written with the SAME cv2 calls as the reference
    (src/Benchmark.cpp:27-32, :133-154), on CV_16U images."""
    cv2 = pytest.importorskip("cv2")
    gtv = cv2.subtract(gt, inp)
    _, gtv = cv2.threshold(gtv, vis_thr, 65536, cv2.THRESH_BINARY)
    _, bgt = cv2.threshold(gt, 1, 65536, cv2.THRESH_BINARY)
    gtv = cv2.subtract(bgt, gtv)
    ev = cv2.subtract(est, inp)
    _, ev = cv2.threshold(ev, vis_thr, 65536, cv2.THRESH_BINARY)
    _, best = cv2.threshold(est, 1, 65536, cv2.THRESH_BINARY)
    ev = cv2.subtract(best, ev)
    best = cv2.bitwise_and(gtv, est)
    ev = cv2.bitwise_or(ev, best)
    inter = cv2.bitwise_and(gtv, ev)
    comb = cv2.bitwise_or(gtv, ev)
    ad = cv2.absdiff(gt, est)
    _, ad = cv2.threshold(ad, err_thr, 65536, cv2.THRESH_BINARY_INV)
    applied = cv2.bitwise_and(inter, ad)
    return np.float32(1) - np.float32(cv2.countNonZero(applied)) / np.float32(cv2.countNonZero(comb)), cv2.countNonZero(applied), cv2.countNonZero(comb)


def test_hodan_error_equals_the_references_cv2_calls():
    """the benchmark's error render path (Benchmark.cpp:18-38) on the headless rasteriser."""
    rng = np.random.default_rng(3)
    verts, faces = torus_mesh()
    tris = [[q[0], q[k], q[k + 1]] for q in faces for k in range(1, len(q) - 1)]
    cam = R.Camera()
    c, s = np.cos(0.3), np.sin(0.3)
    Rg = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    c2, s2 = np.cos(0.36), np.sin(0.36)
    Re = np.array([[c2, 0, s2], [0, 1, 0], [-s2, 0, c2]])
    tg, te = np.array([10.0, -5.0, -700.0]), np.array([14.0, -3.0, -708.0])
    d, _ = R.render_pose(verts, tris, [Rg, Re], [tg, te], cam, colour=False)
    gt, est = d[0], d[1]
    assert gt.max() > 0 and est.max() > 0
    # input depth: the ground-truth render + sensor noise, an occluder in front of part of it, holes
    inp = gt.copy().astype(np.int32)
    inp[gt > 0] += rng.integers(-6, 7, int((gt > 0).sum()))
    inp[gt == 0] = 1200
    inp[200:260, 300:360] = 450
    inp[rng.random(inp.shape) < 0.03] = 0
    inp = np.clip(inp, 0, 65535).astype(np.uint16)
    for vt, et in ((15, 20), (5, 3), (40, 60)):
        want = _hodan_cv2(inp, gt, est, vt, et)
        got = R.hodan_error(inp, gt, est, vt, et)
        assert (got[1], got[2]) == (want[1], want[2]) and got[0] == want[0], (got, want)
    err_same = R.hodan_error(inp, gt, gt)[0]
    assert err_same < R.hodan_error(inp, gt, est)[0] < 1.0
    e2 = R.hodan_error_poses(verts, tris, Rg, tg, Re, te, inp, cam)
    assert e2 == R.hodan_error(inp, gt, est)[0]
