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
preps = [det.prepareBatch(frames, cap=2048*B) for _ in range(2)]
for _ in range(3): det.matchPrepared(preps[0], 80.0)
K = 12
t0 = time.perf_counter()
for _ in range(K): det.matchPrepared(preps[0], 80.0)
print("blocking ms/step %.3f" % ((time.perf_counter()-t0)/K*1e3))
for rep in range(2):
    ts, tc = [], []
    t0 = time.perf_counter(); pending = None
    for k in range(K):
        a = time.perf_counter(); tk = det.submitPrepared(preps[k & 1], 80.0); b = time.perf_counter(); ts.append(b - a)
        if pending is not None:
            det.collectPrepared(*pending); tc.append(time.perf_counter() - b)
        pending = (preps[k & 1], tk)
    det.collectPrepared(*pending)
    dt = time.perf_counter() - t0
    print("pipelined ms/step %.3f  submit avg %.3f ms  collect avg %.3f ms" % (dt/K*1e3, sum(ts)/len(ts)*1e3, sum(tc)/len(tc)*1e3))
# submit only cost (enqueue) when GPU idle
det.synchronize()
a = time.perf_counter(); tk = det.submitPrepared(preps[0], 80.0); b = time.perf_counter(); det.collectPrepared(preps[0], tk); c = time.perf_counter()
print("isolated: submit %.3f ms, collect %.3f ms" % ((b-a)*1e3, (c-b)*1e3))
