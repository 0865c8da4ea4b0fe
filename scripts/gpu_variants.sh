#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f fps  ms/step %.3f'%(d['value'], d['ms_per_step']), ' '.join('%s=%.4f'%(k,v['ms_per_launch']) for k,v in d['kernels'].items()))"; }
echo "== frames 12 (default build)"; timeout 120 python bench.py --steps 20 --warmup 3 --frames 12 --template-cache cache/tpl_cfg2.yml.gz --no-cpu --no-e2e 2>&1 | tail -1 | summ
echo "== frames 1"; timeout 120 python bench.py --steps 50 --warmup 3 --frames 1 --template-cache cache/tpl_cfg2.yml.gz --no-cpu --no-e2e 2>&1 | tail -1 | summ
for v in "-DCW_MINB=7" "-DCW_MINB=8"; do
  LMB200_NVCC_EXTRA="$v" timeout 300 python line_mod_pipeline_b200/build.py -f > /dev/null 2>&1
  echo "== $v"; grep -A3 "similarity_coarse_kernelILb0" line_mod_pipeline_b200/build/build.log | grep -E "registers|spill" | head -2
  timeout 120 python bench.py --steps 10 --warmup 3 --template-cache cache/tpl_cfg2.yml.gz --no-cpu --no-e2e 2>&1 | tail -1 | summ
done
