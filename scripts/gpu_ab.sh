#!/bin/bash
# parity suite on the default build, then the 96-frame bench on the default build and on every cache/liblmb200_*.so variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f fps  e2e %s  ms/step %.3f'%(d['value'], ('%.0f'%d['e2e']['value']) if d.get('e2e') else '-', d['ms_per_step']), ' '.join('%s=%.4f'%(k,v['ms_per_launch']) for k,v in d['kernels'].items()))"; }
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 2>&1 | tail -4
echo "== default, frames 96"; timeout 150 python bench.py --steps 10 --warmup 3 --template-cache cache/tpl_cfg2.yml.gz --no-cpu --no-e2e 2>&1 | tail -1 | summ
for so in cache/liblmb200_*.so; do
  [ -f "$so" ] || continue
  echo "== $so, frames 96"; LMB200_SO=$PWD/$so timeout 150 python bench.py --steps 10 --warmup 3 --template-cache cache/tpl_cfg2.yml.gz --no-cpu --no-e2e 2>&1 | tail -1 | summ
done
echo "== default, frames 12"; timeout 120 python bench.py --steps 20 --warmup 3 --frames 12 --template-cache cache/tpl_cfg2.yml.gz --no-cpu --no-e2e 2>&1 | tail -1 | summ
echo "== default, frames 1"; timeout 120 python bench.py --steps 50 --warmup 3 --frames 1 --template-cache cache/tpl_cfg2.yml.gz --no-cpu --no-e2e 2>&1 | tail -1 | summ
