#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
R=${1:-r01d}
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 1 python __graft_entry__.py --smoke > gpurun_out/memcheck_smoke_${R}.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/memcheck_smoke_${R}.log
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 1 python scripts/exp_memcheck.py > gpurun_out/memcheck_tour_${R}.log 2>&1; echo "memcheck tour rc=$?"; tail -9 gpurun_out/memcheck_tour_${R}.log
