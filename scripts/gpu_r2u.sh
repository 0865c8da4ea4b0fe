#!/bin/bash
# round 2, session u (1 GPU): device-epilogue unit tests, plain and under compute-sanitizer memcheck / racecheck
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
T=${1:-r2u}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "device_epilogue" 2>&1 | tail -15 | tee gpurun_out/${T}_pytest_epilogue.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 -k "device_epilogue" 2>&1 | tail -6 | tee gpurun_out/${T}_memcheck_epilogue.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 -k "device_epilogue and 2-5" 2>&1 | tail -6 | tee gpurun_out/${T}_racecheck_epilogue.log
