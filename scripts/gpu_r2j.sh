#!/bin/bash
# round 2, session j (N GPUs): template-sharded leg with the device-time breakdown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
N=${1:-2}
for F in ${2:-128}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 12 --warmup 3 --only-ts --ts-frames $F > gpurun_out/r2j_ts_N${N}_F$F.log 2> gpurun_out/r2j_ts_N${N}_F$F.err; tail -1 gpurun_out/r2j_ts_N${N}_F$F.log | python -c "
import sys, json
t = json.loads(sys.stdin.read())['template_sharded']; print('TS frames', t['frames_per_step'], 'value', t['value'], 'ms', t['ms_per_step'], 'e2e', t['e2e']['value'], '1gpu', t['full_set_on_1_gpu'], 'eff', t['efficiency_vs_full_set_on_1_gpu'], 'parity', t['parity']); print(t.get('device_ms_per_step'))"
tail -2 gpurun_out/r2j_ts_N${N}_F$F.err
done
