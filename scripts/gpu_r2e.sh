#!/bin/bash
# round 2, session e (2 GPUs): template-sharded parity at configs[3] size (both step variants), bench at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/multi_gpu_worker.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -8 | tee gpurun_out/r2e_multi_gpu_parity_N$N.log
LMB200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2e_bench_N$N.log 2> gpurun_out/r2e_bench_N$N.err; tail -1 gpurun_out/r2e_bench_N$N.log | python -c "
import sys, json
l = json.loads(sys.stdin.read()); print('value', l['value'], 'e2e', l['e2e']['value'], l['e2e']['frac_of_h2d_roof'], 'roofs', l['roofs']); print(json.dumps(l['template_sharded'], indent=1)); print(l['config']['cpu_affinity'])"
grep "allgather fetch" gpurun_out/r2e_bench_N$N.err | tail -4
