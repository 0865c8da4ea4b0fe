#!/bin/bash
# round 2, session d: parity with the tuned frame-side kernels, the reworked bench (all legs) on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 | tee gpurun_out/r2d_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench.log 2>&1; tail -c 6000 gpurun_out/r2d_bench.log
