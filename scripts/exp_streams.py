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
prep = det.prepareBatch(frames, cap=2048*B)
for _ in range(3): det.matchPrepared(prep, 80.0)
def run(tag):
    det.matchPrepared(prep, 80.0)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); det.matchPrepared(prep, 80.0); ts.append((time.perf_counter()-t0)*1e3)
    ts.sort(); print(tag, "ms/step median %.3f min %.3f max %.3f" % (ts[len(ts)//2], ts[0], ts[-1]), flush=True)
run("default")
for nx in (3, 4):
    for G, ch in ((6, 12), (6, 16), (4, 24), (4, 16), (3, 32), (8, 12), (8, 8)):
        os.environ["LMB200_XSTREAMS"] = str(nx); os.environ["LMB200_GROUPS"] = str(G); os.environ["LMB200_CHUNK"] = str(ch)
        run("xstreams %d groups %d chunk %d" % (nx, G, ch))
