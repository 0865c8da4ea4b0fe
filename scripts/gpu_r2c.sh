#!/bin/bash
# round 2, session c: parity after the DN chain fix + ncu --set full of the rebuilt frame-side kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 | tee gpurun_out/r2c_pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pyrdown_planar|cg_quantize2|dn_median|spread_strip|spread_flat' -s 14 -c 7 -o gpurun_out/r2c_frame python bench.py --steps 1 --warmup 3 --frames 96 --no-e2e --no-cpu > gpurun_out/r2c_ncu.log 2>&1; tail -2 gpurun_out/r2c_ncu.log
