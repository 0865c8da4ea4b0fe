#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for v in "-DCW_MINB=8 -DCW_GROUP=3" "-DCW_MINB=6 -DCW_GROUP=3" "-DCW_MINB=8 -DCW_GROUP=2" "-DCW_MINB=10 -DCW_GROUP=2" "-DCW_MINB=12 -DCW_GROUP=1"; do
  LMB200_NVCC_EXTRA="$v" python line_mod_pipeline_b200/build.py -f > /dev/null 2>&1
  echo "== $v"; grep -A3 "similarity_coarse_kernelILb0" line_mod_pipeline_b200/build/build.log | grep -E "registers|spill" | head -2
  timeout 300 python bench.py --steps 10 --warmup 3 --template-cache cache/tpl_cfg2.yml.gz --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f fps'%d['value'], ' '.join('%s=%.3f'%(k,v['ms_per_launch']) for k,v in d['kernels'].items() if k in ('sim_coarse','sim_local','linearize','pack')), 'matches', d['matches_per_step'])"
done
