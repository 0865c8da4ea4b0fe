#!/bin/bash
# first GPU session: parity tests, smoke, short bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --frames 32 > gpurun_out/bench1.log 2>&1; tail -5 gpurun_out/bench1.log
