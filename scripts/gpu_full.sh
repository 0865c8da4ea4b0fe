#!/bin/bash
# full single-GPU session: all GPU tests, smoke, bench, reference arm, ncu launch list + full capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
R=${1:-r01b}
timeout 900 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_${R}.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 --template-cache cache/tpl_cfg2.yml.gz > gpurun_out/bench_${R}.log 2>&1; tail -1 gpurun_out/bench_${R}.log
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${R}.log 2>&1; tail -1 gpurun_out/bench_ref_${R}.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 3 --frames 96 --no-e2e --no-cpu --template-cache cache/tpl_cfg2.yml.gz > gpurun_out/ncu_launch_${R}.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'similarity_coarse|similarity_local|spread_linearize|cg_quantize|dn_quantize|median5|pyrdown|pack_' -s 16 -c 16 -o gpurun_out/prof_${R} python bench.py --steps 1 --warmup 3 --frames 96 --no-e2e --no-cpu --template-cache cache/tpl_cfg2.yml.gz > gpurun_out/ncu_full_${R}.log 2>&1; tail -1 gpurun_out/ncu_full_${R}.log
