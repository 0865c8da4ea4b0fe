#!/bin/bash
# 2-GPU check: template-sharded NCCL parity, then the frames-mode and templates-mode bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}; R=${2:-r01d}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_worker.py 2>&1 | grep -E "MULTI_GPU|DIFFER|Error|error" | head -5 | tee gpurun_out/multi_parity_N${N}_$R.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --template-cache cache/tpl_cfg2.yml.gz --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_frames_N${N}_$R.log | cut -c1-700
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 --shard templates 2>&1 | tail -1 | tee gpurun_out/bench_templates_N${N}_$R.log | cut -c1-400
