#!/bin/bash
# round 2, final multi-GPU sanity (2 GPUs): the parity worker (20 000 templates, every fetch path) on the final code
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
N=${1:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/multi_gpu_worker.py > gpurun_out/r02e_multi_gpu_parity_N$N.full 2>&1
grep "MULTI_GPU\|DIFFER\|disagree\|rror\|differs" gpurun_out/r02e_multi_gpu_parity_N$N.full | tail -8 | tee gpurun_out/r02e_multi_gpu_parity_N$N.log
