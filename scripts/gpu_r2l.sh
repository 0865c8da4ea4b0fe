#!/bin/bash
# round 2, session l (N GPUs): parity worker + template-sharded leg with the epilogue thread, lane split and prepared host calls
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
N=${1:-2}
T=${2:-r2l}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/multi_gpu_worker.py > gpurun_out/${T}_multi_gpu_parity_N$N.full 2>&1
grep "MULTI_GPU\|DIFFER\|disagree\|rror\|differs" gpurun_out/${T}_multi_gpu_parity_N$N.full | tail -8 | tee gpurun_out/${T}_multi_gpu_parity_N$N.log
LMB200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 20 --warmup 3 --only-ts > gpurun_out/${T}_ts_N$N.log 2> gpurun_out/${T}_ts_N$N.err; tail -1 gpurun_out/${T}_ts_N$N.log | python -c "
import sys, json
t = json.loads(sys.stdin.read())['template_sharded']; print('TS frames', t['frames_per_step'], 'value', t['value'], 'ms', t['ms_per_step'], 'e2e', t['e2e']['value'], '1gpu', t['full_set_on_1_gpu'], 'eff', t['efficiency_vs_full_set_on_1_gpu'], 'parity', t['parity']); print(t.get('device_ms_per_step'))"
grep "allgather fetch" gpurun_out/${T}_ts_N$N.err | tail -4
grep "sharded submit" gpurun_out/${T}_ts_N$N.err | tail -4
grep -i "error\|Traceback" gpurun_out/${T}_ts_N$N.err | head -5
