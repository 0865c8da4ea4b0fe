#!/bin/bash
# round 2, session h: full parity (CUDA-graph single-frame path, post-match colour check, drop-in TU on GPU), bench N=1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 | tee gpurun_out/r2h_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_bench.log 2>&1; tail -1 gpurun_out/r2h_bench.log | python -c "
import sys, json
l = json.loads(sys.stdin.read()); print('value', l['value'], 'e2e', l['e2e']['value'], l['e2e']['frac_of_h2d_roof'], 'single', l['single_frame']); t=l['template_sharded']; print('TS', t['value'], t['e2e']['value'], t['parity']); print({k:v['ms_per_launch'] for k,v in l['kernels'].items()})"
tail -3 gpurun_out/r2h_bench.log | head -2 | cut -c1-300
