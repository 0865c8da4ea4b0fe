import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import line_mod_pipeline_b200 as lm
from line_mod_pipeline_b200 import synth
B = 96
det0 = lm.Detector.readCache("cache/tpl_cfg2.lmb200") if os.path.exists("cache/tpl_cfg2.lmb200") else lm.Detector.read("cache/tpl_cfg2.yml.gz")
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
for nx in (1, 2, 3, 4):
    for G in (3, 4, 6):
        for ch in (8, 12, 16):
            if ch > B // G: continue
            os.environ["LMB200_XSTREAMS"] = str(nx); os.environ["LMB200_GROUPS"] = str(G); os.environ["LMB200_CHUNK"] = str(ch)
            det.matchPrepared(prep, 80.0)
            t0 = time.perf_counter()
            for _ in range(5): det.matchPrepared(prep, 80.0)
            print("xstreams", nx, "groups", G, "chunk", ch, "ms/step %.3f" % ((time.perf_counter()-t0)/5*1e3), flush=True)
