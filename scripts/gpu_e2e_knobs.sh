#!/bin/bash
# e2e (host frames -> matches) under different chunk schedules of the batch pipeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('value %.0f  e2e %.0f (blocking %s)  ms/step %.3f'%(d['value'], e['value'], e.get('blocking_call'), d['ms_per_step']))"; }
run() { echo "== $*"; env "$@" timeout 150 python bench.py --steps 10 --warmup 3 --template-cache cache/tpl_cfg2.yml.gz --no-cpu 2>&1 | tail -1 | summ; }
run LMB200_GROUPS=6 LMB200_CHUNK=12
run LMB200_GROUPS=4 LMB200_CHUNK=24
run LMB200_GROUPS=3 LMB200_CHUNK=32
run LMB200_GROUPS=4 LMB200_CHUNK=16
run LMB200_GROUPS=3 LMB200_CHUNK=32 LMB200_XSTREAMS=2
