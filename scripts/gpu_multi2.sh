#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_worker.py 2>&1 | grep -E "MULTI_GPU|DIFFER|Error|error" | head -5 | tee gpurun_out/multi_parity_N$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 --shard templates 2>&1 | tail -1 | tee gpurun_out/bench_templates_N$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('templates mode: N', d['n_gpus'], 'value %.0f fps  ms/step %.3f'%(d['value'], d['ms_per_step']), d['config']['templates'], 'templates', ' '.join('%s=%.3f'%(k,v['ms_per_launch']) for k,v in d['kernels'].items()))"
