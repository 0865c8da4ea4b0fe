#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'similarity_coarse|similarity_local' -s 6 -c 2 -o gpurun_out/prof_coarse python bench.py --steps 1 --warmup 3 --frames 96 --no-e2e --no-cpu --template-cache cache/tpl_cfg2.yml.gz > gpurun_out/ncu_coarse.log 2>&1; tail -2 gpurun_out/ncu_coarse.log
