#!/bin/bash
# round 2, session i (8 GPUs): template-sharded leg (128 frames/step, 3 slot groups, fetch communicator) + multi-GPU parity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/multi_gpu_worker.py 2>&1 | grep "MULTI_GPU\|DIFFER\|disagree\|rror" | tail -5 | tee gpurun_out/r2i_multi_gpu_parity_N$N.log
for F in 128 64; do
LMB200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 12 --warmup 3 --only-ts --ts-frames $F > gpurun_out/r2i_ts_N${N}_F$F.log 2> gpurun_out/r2i_ts_N${N}_F$F.err; tail -1 gpurun_out/r2i_ts_N${N}_F$F.log | python -c "
import sys, json
t = json.loads(sys.stdin.read())['template_sharded']; print('TS frames', t['frames_per_step'], 'value', t['value'], 'ms', t['ms_per_step'], 'e2e', t['e2e']['value'], '1gpu', t['full_set_on_1_gpu'], 'eff', t['efficiency_vs_full_set_on_1_gpu'], 'parity', t['parity'])"
grep "allgather fetch" gpurun_out/r2i_ts_N${N}_F$F.err | tail -6
done
