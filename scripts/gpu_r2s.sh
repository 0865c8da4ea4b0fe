#!/bin/bash
# round 2, session s (N GPUs): the driver's own command line — full bench (frame-sharded main leg + template-sharded leg) and the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
N=${1:-2}
T=${2:-r2s}
SECONDS=0; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${T}_bench_N$N.log 2> gpurun_out/${T}_bench_N$N.err
echo "bench wall ${SECONDS}s"
tail -1 gpurun_out/${T}_bench_N$N.log > gpurun_out/${T}_bench_N${N}_line.json
python - <<PY
import json
t = json.load(open("gpurun_out/${T}_bench_N${N}_line.json"))
print('value', t['value'], 'ms', t['ms_per_step'], 'e2e', t['e2e']['value'], t['e2e'].get('frac_of_h2d_roof'), 'launches', t['gpu_launches'])
ts = t.get('template_sharded') or {}
print('TS value', ts.get('value'), 'ms', ts.get('ms_per_step'), 'e2e', (ts.get('e2e') or {}).get('value'), 'eff', ts.get('efficiency_vs_full_set_on_1_gpu'), 'parity', ts.get('parity'), ts.get('variants'))
print(t.get('roofs'))
PY
grep -i "error\|Traceback" gpurun_out/${T}_bench_N$N.err | head -5
