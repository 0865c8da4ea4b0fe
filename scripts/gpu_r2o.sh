#!/bin/bash
# round 2, session o (N GPUs): template-sharded leg only (in-run parity bit against the oracle), spread on lane 3, 4 slot groups
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
N=${1:-8}
T=${2:-r2o}
LMB200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 30 --warmup 3 --only-ts > gpurun_out/${T}_ts_N$N.log 2> gpurun_out/${T}_ts_N$N.err; tail -1 gpurun_out/${T}_ts_N$N.log | python -c "
import sys, json
t = json.loads(sys.stdin.read())['template_sharded']; print('TS frames', t['frames_per_step'], 'value', t['value'], 'ms', t['ms_per_step'], 'e2e', t['e2e']['value'], '1gpu', t['full_set_on_1_gpu'], 'eff', t['efficiency_vs_full_set_on_1_gpu'], 'parity', t['parity']); print(t.get('device_ms_per_step')); print(t.get('variants'))"
grep "allgather fetch" gpurun_out/${T}_ts_N$N.err | awk '{print $6}' | head -60 | tr '\n' ' '
grep -i "error\|Traceback" gpurun_out/${T}_ts_N$N.err | head -5
