#!/bin/bash
# round 2, session r (1 GPU): e2e of the streaming batch path with more frame slots / other chunk schedules
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('value %.0f  e2e %.0f frac %.3f (blocking %.0f)  ms/step %.3f'%(d['value'], e['value'], e['frac_of_h2d_roof'], e['blocking_call']['value'], d['ms_per_step']))"; }
run() { echo "== $*"; env "${@:2}" timeout 150 python bench.py --steps 16 --warmup 3 --no-cpu --no-ts --no-extra --slots $1 2>&1 | tail -1 | summ; }
run 64 LMB200_GROUPS=4 LMB200_CHUNK=16
run 128 LMB200_GROUPS=4 LMB200_CHUNK=32
run 128 LMB200_GROUPS=4 LMB200_CHUNK=22
run 128 LMB200_GROUPS=5 LMB200_CHUNK=24 LMB200_XSTREAMS=2
run 192 LMB200_GROUPS=3 LMB200_CHUNK=64 LMB200_XSTREAMS=2
