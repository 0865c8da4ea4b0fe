#!/bin/bash
# round 2, final sanity (1 GPU): full GPU suite, smoke, default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
T=${1:-r02e}
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4 | tee gpurun_out/${T}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.log 2>&1; tail -1 gpurun_out/${T}_bench.log > gpurun_out/${T}_bench_line.json; python - <<PY
import json
t = json.load(open("gpurun_out/${T}_bench_line.json"))
print('value', t['value'], 'ms', t['ms_per_step'], 'steps', t['steps'], 'serial', t['serial_profiled_pass']['value'], 'e2e', t['e2e']['value'], 'single', t['single_frame']['median_ms'], 'frac', t['roofline']['frac'], 'cpu', t['cpu_baseline']['value'])
PY
