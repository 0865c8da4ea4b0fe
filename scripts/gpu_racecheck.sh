#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
R=${1:-r01d}
timeout 160 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 1 python scripts/exp_memcheck.py > gpurun_out/racecheck_tour_${R}.log 2>&1; echo "racecheck tour rc=$?"; grep -c "Race reported\|hazard" gpurun_out/racecheck_tour_${R}.log; tail -12 gpurun_out/racecheck_tour_${R}.log | cut -c1-220
