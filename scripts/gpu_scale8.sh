#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
(timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --template-cache cache/tpl_cfg2.yml.gz 2>&1 | tail -1 | tee gpurun_out/bench_frames_N$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('frames mode: N', d['n_gpus'], 'value %.0f fps'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'blocking %.0f'%d['e2e']['blocking_call']['value'])") 
(timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 --shard templates --templates 2500 2>&1 | tail -1 | tee gpurun_out/bench_templates_N$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('templates mode (config 4: %d templates): N'%d['config']['templates'], d['n_gpus'], 'value %.0f fps  ms/step %.3f'%(d['value'], d['ms_per_step']), 'matches', d['matches_per_step'])")
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_worker.py 2>&1 | grep -E "MULTI_GPU|DIFFER" | head -3 | tee gpurun_out/multi_parity_N$N.log
