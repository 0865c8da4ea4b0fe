"""Times template creation: one lmb200_add_template per view vs lmb200_add_templates."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from line_mod_pipeline_b200 import Detector, ColorGradient, DepthNormal, synth

views, masks = [], []
for i in range(8):
    bgr, depth = synth.make_frame(30 + i)
    for m in synth.object_masks(30 + i)[:16] + synth.planted_masks(16, seed=60 + i):
        views.append([bgr, depth]); masks.append(m)
n = len(views)
for name, fn in (("single", 0), ("bulk", 1), ("single", 0), ("bulk", 1)):
    det = Detector([ColorGradient(), DepthNormal()], [5, 8])
    det.addTemplate(views[0], "warm", masks[0])
    t0 = time.perf_counter()
    if fn:
        res = det.addTemplates(views, "obj", masks)
    else:
        res = [det.addTemplate(v, "obj", m) for v, m in zip(views, masks)]
    dt = time.perf_counter() - t0
    print("%s: %d views, %d templates, %.1f ms total, %.3f ms/view, %.0f views/s" %
          (name, n, sum(1 for t, _ in res if t >= 0), dt * 1e3, dt * 1e3 / n, n / dt), flush=True)
