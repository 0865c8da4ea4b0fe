#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_worker.py 2>&1 | tail -15 | tee gpurun_out/multi_parity_N$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --template-cache cache/tpl_cfg2.yml.gz 2>&1 | tail -3 | tee gpurun_out/bench_frames_N$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 --shard templates 2>&1 | tail -3 | tee gpurun_out/bench_templates_N$N.log
timeout 300 python bench.py --steps 10 --warmup 3 --template-cache cache/tpl_cfg2.yml.gz 2>&1 | tail -1 | tee gpurun_out/bench_N1.log
