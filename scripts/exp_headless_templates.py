"""Headless template generation end to end (SURVEY 8f-4 + 8f-1 + the hot path): render a viewpoint sweep of a
procedural model on the host threads, turn the views into templates with the bulk addTemplates, then find one of the
rendered views again with match()."""
import sys, time
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import line_mod_pipeline_b200 as lm
from line_mod_pipeline_b200 import render as R
from test_render_host import torus_mesh

verts, faces = torus_mesh(R0=60.0, r0=22.0)
tris = []
for q in faces:
    for k in range(1, len(q) - 1):
        tris.append((q[0], q[k], q[k + 1]))
tris = np.array(tris, np.int32)
eyes = []
for radius in (450.0, 550.0, 650.0, 750.0):
    for el in np.linspace(0.15, 1.45, 8):
        for az in np.linspace(0.0, 2 * np.pi, 12, endpoint=False):
            eyes.append((radius * np.cos(el) * np.sin(az), radius * np.sin(el), radius * np.cos(el) * np.cos(az)))
t0 = time.perf_counter()
depth, colour = R.render_lookat(verts, tris, eyes)
t_render = time.perf_counter() - t0
masks = [np.where(d > 0, 255, 0).astype(np.uint8) for d in depth]
# shade the silhouette a little so that the colour modality has gradients inside the outline too
shaded = [np.where(c > 0, (80 + (d.astype(np.int32) % 64) * 2)[..., None], 0).astype(np.uint8) for c, d in zip(colour, depth)]
det = lm.getDefaultLINEMOD()
views = [[shaded[i], depth[i]] for i in range(len(eyes))]
t0 = time.perf_counter()
res = det.addTemplates(views, "torus", masks)
t_add = time.perf_counter() - t0
ok = [i for i, (tid, _) in enumerate(res) if tid >= 0]
print("views %d: render %.2f ms/view, addTemplates %.2f ms/view -> %d templates" %
      (len(eyes), 1e3 * t_render / len(eyes), 1e3 * t_add / len(eyes), len(ok)), flush=True)
probe = ok[len(ok) // 2]
m = det.match(views[probe], 90.0)
best = m[0] if len(m) else None
print("scene = view %d (template %d): %d matches, best %s" % (probe, res[probe][0], len(m), best), flush=True)
assert best is not None and int(best["template_id"]) == res[probe][0] and float(best["similarity"]) >= 99.0
print("HEADLESS_OK")
