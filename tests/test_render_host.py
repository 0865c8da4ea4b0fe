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
