#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q "$@" 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
tail -120 gpurun_out/pytest_gpu.log
