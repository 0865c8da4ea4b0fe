#!/bin/bash
# round 2, session w (1 GPU): A/B of the coarse kernel's neighbour-chunk SHFL variant (-DCW_SHFL build loaded through LMB200_SO)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
T=${1:-r2w}
summ() { python -c "
import sys, json
t = json.loads(sys.stdin.read()); print('value', round(t['value']), 'ms', round(t['ms_per_step'], 4), 'coarse', t['kernels']['sim_coarse']['ms_per_launch'], 'local', t['kernels']['sim_local']['ms_per_launch'])"; }
LMB200_SO=$PWD/line_mod_pipeline_b200/liblmb200_shfl.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "match_parity or config2_full or config4 or similarity_maps_config2 or wide_sums or degenerate" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_shfl.log
for v in base shfl base shfl; do
  if [ $v = shfl ]; then export LMB200_SO=$PWD/line_mod_pipeline_b200/liblmb200_shfl.so; else unset LMB200_SO; fi
  echo "== $v"; timeout 300 python bench.py --steps 20 --warmup 3 --no-ts --no-extra --no-cpu --no-e2e 2>/dev/null | tail -1 | summ
done 2>&1 | tee gpurun_out/${T}_ab.log
