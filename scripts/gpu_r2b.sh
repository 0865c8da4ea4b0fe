#!/bin/bash
# round 2, session b: rebuilt frame-side kernels — parity (frame-side tests first), memcheck, A/B against round 1's kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "frame_side or masks or generic" 2>&1 | tail -25 | tee gpurun_out/r2b_pytest_frame.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25 | tee gpurun_out/r2b_pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 -k "frame_side_fixture or ragged or uncovered" 2>&1 | tail -12 | tee gpurun_out/r2b_memcheck.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2b_bench.log 2>&1; tail -1 gpurun_out/r2b_bench.log | python -c "
import sys, json
l = json.loads(sys.stdin.read()); print('NEW value', l['value'], 'e2e', l['e2e']['value'], 'ms', l['ms_per_step']); print({k: v['ms_per_launch'] for k, v in l['kernels'].items()}); print({k: v['launches'] for k, v in l['kernels'].items()})"
LMB200_GENERIC_FRAME=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2b_bench_generic.log 2>&1; tail -1 gpurun_out/r2b_bench_generic.log | python -c "
import sys, json
l = json.loads(sys.stdin.read()); print('OLD value', l['value'], 'e2e', l['e2e']['value'], 'ms', l['ms_per_step']); print({k: v['ms_per_launch'] for k, v in l['kernels'].items()})"
