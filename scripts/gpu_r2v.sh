#!/bin/bash
# round 2, session v (1 GPU): parity tests + short bench after the decimated fast path of the spread kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
T=${1:-r2v}
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4 | tee gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-ts --no-extra --no-cpu > gpurun_out/${T}_bench.log 2>&1; tail -1 gpurun_out/${T}_bench.log | python -c "
import sys, json
t = json.loads(sys.stdin.read()); print('value', t['value'], 'ms', t['ms_per_step'], 'e2e', t['e2e']['value'], 'single', t.get('single_frame', {}).get('median_ms'))
for k, v in t['kernels'].items(): print(k, v.get('ms_per_launch'), v.get('launches'), v.get('frac_of_hbm'))"
