#!/bin/bash
# round 2, session t (N GPUs): template-sharded leg incl. the 2-D (template shards x frame groups) layout
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
N=${1:-4}
T=${2:-r2t}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 20 --warmup 3 --only-ts > gpurun_out/${T}_ts_N$N.log 2> gpurun_out/${T}_ts_N$N.err; tail -1 gpurun_out/${T}_ts_N$N.log | python -c "
import sys, json
t = json.loads(sys.stdin.read())['template_sharded']; print('TS frames', t['frames_per_step'], 'value', t['value'], 'ms', t['ms_per_step'], 'e2e', t['e2e']['value'], '1gpu', t['full_set_on_1_gpu'], 'eff', t['efficiency_vs_full_set_on_1_gpu'], 'parity', t['parity']); print(t.get('device_ms_per_step')); print(t.get('variants')); print(t.get('grid_2d'))"
grep -i "error\|Traceback" -A5 gpurun_out/${T}_ts_N$N.err | head -20
