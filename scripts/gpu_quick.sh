#!/bin/bash
# quick regression + perf loop: parity subset, then bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 2>&1 | tail -15
timeout 150 python bench.py --steps 10 --warmup 3 --template-cache cache/tpl_cfg2.yml.gz --no-cpu 2>&1 | tail -1 > gpurun_out/bench_quick.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.log').read())
print("value %.0f fps  e2e %.0f fps  ms/step %.3f" % (d['value'], d['e2e']['value'] if d['e2e'] else 0, d['ms_per_step']))
for k,v in d['kernels'].items(): print("  %-12s %8.4f ms/launch x%d  %s" % (k, v['ms_per_launch'], v['launches'], ("%.0f GB/s"%v['alg_GBps']) if 'alg_GBps' in v else ''))
print(d['roofline'])
PY
