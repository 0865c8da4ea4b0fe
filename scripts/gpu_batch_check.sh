#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "batch or pipelined or ragged or overflow or resident" 2>&1 | tail -3
timeout 150 python bench.py --steps 10 --warmup 3 --template-cache cache/tpl_cfg2.yml.gz --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_batch_check.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('value %.0f  e2e %.0f (blocking %s)  ms/step %.3f'%(d['value'], e['value'], e.get('blocking_call'), d['ms_per_step']))"
