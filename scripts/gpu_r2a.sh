#!/bin/bash
# round 2, session a: new parity tests (similarity maps, config 5 at 10 000, stale resident ranges), measured roofs, baseline bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 | tee gpurun_out/r2a_pytest_gpu.log
python - <<'PY' 2>&1 | tee gpurun_out/r2a_microbench.log
import ctypes as C
import line_mod_pipeline_b200 as lm
L = lm.capi.lib()
for name, kind in (("L2_READ", 0), ("L1_READ", 1), ("HBM_READ", 2), ("H2D", 3)):
    for rep in range(2):
        g = C.c_double(0)
        rc = L.lmb200_microbench(kind, 0, 10, C.byref(g))
        print("microbench %-8s rc=%d %.1f GB/s" % (name, rc, g.value))
for mb in (8, 24, 48, 96, 160):
    g = C.c_double(0); L.lmb200_microbench(0, mb << 20, 10, C.byref(g)); print("L2_READ %d MB: %.1f GB/s" % (mb, g.value))
for kb in (4, 8, 16, 24, 32):
    g = C.c_double(0); L.lmb200_microbench(1, kb << 10, 10, C.byref(g)); print("L1_READ %d KB per CTA x 8 CTAs/SM: %.1f GB/s" % (kb, g.value))
PY
timeout 300 python bench.py --steps 20 --warmup 3 --template-cache cache/tpl_cfg2.yml.gz > gpurun_out/r2a_bench.log 2>&1; tail -1 gpurun_out/r2a_bench.log
