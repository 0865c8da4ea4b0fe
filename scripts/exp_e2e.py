import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import line_mod_pipeline_b200 as lm
from line_mod_pipeline_b200 import synth
B = 96
det0 = lm.Detector.read("cache/tpl_cfg2.yml.gz")
det = lm.getDefaultLINEMOD(max_batch=B)
for cid in det0.classIds():
    for t in range(det0.numTemplates(cid)):
        det.addSyntheticTemplate(det0.getTemplates(cid, t), cid)
L = lm.capi.lib()
nb, nd = 480*640*3, 480*640*2
ptr = C.c_void_p(); L.lmb200_host_alloc(B*(nb+nd), C.byref(ptr))
host = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(B*(nb+nd),))
frames = []
for i in range(B):
    bgr, depth = synth.make_frame(i % 8)
    hb = host[i*(nb+nd): i*(nb+nd)+nb].reshape(480,640,3); hd = host[i*(nb+nd)+nb:(i+1)*(nb+nd)].view(np.uint16).reshape(480,640)
    hb[:] = bgr; hd[:] = depth; frames.append([hb, hd])
det.uploadFrames(frames, 0)
def t(fn, n=5):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
print("upload only ms", t(lambda: det.uploadFrames(frames, 0)), "-> GB/s", B*(nb+nd)/1e6/ t(lambda: det.uploadFrames(frames, 0)))
def res():
    det.matchResident(0, B, 80.0); det.synchronize()
print("resident ms", t(res))
def resf():
    det.matchResident(0, B, 80.0); det.fetchResident(0, B, cap=16384*B)
print("resident+fetch ms", t(resf))
for ch in (1, 2, 4, 8, 16, 24, 48):
    os.environ["LMB200_CHUNK"] = str(ch)
    print("batch chunk", ch, "ms", t(lambda: det.matchBatch(frames, 80.0, cap=16384*B)))
