#!/bin/bash
# round 2, session x (1 GPU): resident-path tests + full default bench after the frame-lane overlap of lmb200_match_resident
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
T=${1:-r2x}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "resident or config3 or pipelined or overflow or match_parity or masks" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
SECONDS=0
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.log 2>&1; echo "bench wall ${SECONDS}s"
tail -1 gpurun_out/${T}_bench.log > gpurun_out/${T}_bench_line.json; python - <<PY
import json
t = json.load(open("gpurun_out/${T}_bench_line.json"))
print('value', t['value'], 'ms', t['ms_per_step'], 'serial', t['serial_profiled_pass']['value'], t['serial_profiled_pass']['ms_per_step'], 'e2e', t['e2e']['value'], 'single', t['single_frame']['median_ms'])
print('roofline', t['roofline']['achieved'], t['roofline']['peak'], t['roofline']['frac'], 'clocks', t['clocks'])
print('ts 1gpu', t['template_sharded']['value'], t['template_sharded']['parity'])
PY
grep -i "error\|Traceback\|assert" gpurun_out/${T}_bench.log | head -5
