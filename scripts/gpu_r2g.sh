#!/bin/bash
# round 2, session g (N GPUs): template-sharded leg only, with the fetch trace
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
N=${1:-8}
LMB200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 10 --warmup 3 --only-ts > gpurun_out/r2g_ts_N$N.log 2> gpurun_out/r2g_ts_N$N.err; tail -1 gpurun_out/r2g_ts_N$N.log | python -c "
import sys, json
t = json.loads(sys.stdin.read())['template_sharded']; print('TS value', t['value'], 'ms', t['ms_per_step'], 'e2e', t['e2e']['value'], '1gpu', t['full_set_on_1_gpu'], 'eff', t['efficiency_vs_full_set_on_1_gpu'], 'parity', t['parity'])"
grep "allgather fetch" gpurun_out/r2g_ts_N$N.err | tail -30
