#!/bin/bash
# round 2, session f (N GPUs): multi-GPU parity, bench at N (no trace: the trace mode serialises the streams)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/multi_gpu_worker.py 2>&1 | grep "MULTI_GPU\|DIFFER\|disagree\|Error\|error" | tail -8 | tee gpurun_out/r2f_multi_gpu_parity_N$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2f_bench_N$N.log 2> gpurun_out/r2f_bench_N$N.err; tail -1 gpurun_out/r2f_bench_N$N.log | python -c "
import sys, json
l = json.loads(sys.stdin.read()); print('value', l['value'], 'e2e', l['e2e']['value'], l['e2e']['frac_of_h2d_roof'], 'h2d roof', l['roofs']['h2d_GBps_all_gpus']); t=l['template_sharded']; print('TS value', t['value'], 'ms', t['ms_per_step'], 'e2e', t['e2e']['value'], '1gpu', t['full_set_on_1_gpu'], 'eff', t['efficiency_vs_full_set_on_1_gpu'], 'parity', t['parity'])"
tail -3 gpurun_out/r2f_bench_N$N.err
