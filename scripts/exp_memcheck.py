"""Small tour of the product's kernels for compute-sanitizer (memory safety only; parity is pytest's job)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import line_mod_pipeline_b200 as lm
from line_mod_pipeline_b200 import synth

def tour(name, mods, T, rows, cols, levels, n_rand, thr):
    det = lm.Detector([lm.ColorGradient() if m == "cg" else lm.DepthNormal() for m in mods], list(T), max_batch=8)
    bgr, depth = synth.make_frame(1, rows, cols, n_shapes=12)
    src = [bgr if m == "cg" else depth for m in mods]
    masks = synth.object_masks(1, rows, cols, n_shapes=12)[:6]
    res = det.addTemplates([src] * len(masks), "obj", masks) if masks else []
    for tp in synth.random_templates(n_rand, len(mods), levels, seed=3):
        det.addSyntheticTemplate(tp, "rand")
    m1 = det.match(src, thr)
    frames = [[f[0] if m == "cg" else f[1] for m in mods] for f in (synth.make_frame(i, rows, cols, n_shapes=12) for i in range(5))]
    mb = det.matchBatch(frames, thr)
    print(name, "templates", det.numTemplates(), "planted ok", sum(1 for t, _ in res if t >= 0), "matches", len(m1), [len(b) for b in mb], flush=True)
    det.close()

tour("cfgA", ("cg", "dn"), (5, 8), 480, 640, 2, 60, 60.0)
tour("three-level", ("cg", "dn"), (4, 8, 16), 384, 512, 3, 60, 60.0)
tour("single-level", ("cg", "dn"), (8,), 240, 320, 1, 60, 45.0)
tour("colour-only", ("cg",), (2, 8), 240, 320, 2, 60, 60.0)
# odd-sized template images: the unaligned (byte-wise) staging paths of pyrDown / cg_quantize
det = lm.getDefaultLINEMOD()
bgr, depth = synth.make_frame(2)
ob, od = np.ascontiguousarray(bgr[:251, :333]), np.ascontiguousarray(depth[:251, :333])
mask = np.zeros((251, 333), np.uint8); mask[40:200, 60:280] = 255
print("odd-size addTemplate:", det.addTemplate([ob, od], "odd", mask), det.addTemplates([[ob, od]] * 3, "odd", [mask] * 3), flush=True)
