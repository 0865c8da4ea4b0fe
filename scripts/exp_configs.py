"""Latency / throughput of the other BASELINE configs on one GPU (product only; parity for them is in tests/)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import line_mod_pipeline_b200 as lm
from line_mod_pipeline_b200 import synth

def timed(fn, reps=12):
    fn(); fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); t.append(time.perf_counter() - t0)
    t.sort()
    return 1e3 * t[len(t) // 2], r

# config 5: 1920x1088, T={4,8,16}, 10 000 templates
rows, cols = 1088, 1920
def frame5(i):
    b, d = synth.make_frame(i, 1080, cols, n_shapes=60)
    return [np.concatenate([b, np.zeros((8, cols, 3), np.uint8)], 0), np.concatenate([d, np.zeros((8, cols), np.uint16)], 0)]
det = lm.Detector([lm.ColorGradient(), lm.DepthNormal()], [4, 8, 16], max_batch=8)
f0 = frame5(7)
masks = [np.concatenate([m, np.zeros((8, cols), np.uint8)], 0) for m in synth.object_masks(7, 1080, cols, n_shapes=60, min_px=4000)][:12]
res = det.addTemplates([f0] * len(masks), "planted", masks)
n = sum(1 for t, _ in res if t >= 0)
for tp in synth.random_templates(10000 - n, 2, 3, seed=99, wh_range=(80, 300)):
    det.addSyntheticTemplate(tp, "rand")
ms, m = timed(lambda: det.match(f0, 80.0))
print("config 5: 1920x1088 x %d templates (3 levels): single frame %.3f ms, %d matches" % (det.numTemplates(), ms, len(m)), flush=True)
frames = [frame5(i) for i in range(8)]
ms, b = timed(lambda: det.matchBatch(frames, 80.0), reps=6)
print("config 5: batch of 8 frames %.3f ms -> %.0f frames/s" % (ms, 8e3 / ms), flush=True)
det.close()

# config 4 on one GPU: 640x480 x 20 000 templates
det = lm.getDefaultLINEMOD(max_batch=32)
g0 = list(synth.make_frame(0))
masks = synth.object_masks(0) + synth.planted_masks(200, seed=17)
res = det.addTemplates([g0] * len(masks), "planted", masks)
n = sum(1 for t, _ in res if t >= 0)
for tp in synth.random_templates(20000 - n, 2, 2, seed=99):
    det.addSyntheticTemplate(tp, "rand")
ms, m = timed(lambda: det.match(g0, 80.0))
print("config 4 (one GPU): 640x480 x %d templates: single frame %.3f ms, %d matches" % (det.numTemplates(), ms, len(m)), flush=True)
frames = [list(synth.make_frame(i)) for i in range(32)]
ms, b = timed(lambda: det.matchBatch(frames, 80.0), reps=6)
print("config 4 (one GPU): batch of 32 frames %.3f ms -> %.0f frames/s" % (ms, 32e3 / ms), flush=True)
