#!/bin/bash
# sanitizer + ncu launch list + full capture of the similarity kernels (B200_PROFILING.md recipe)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
R=${1:-r01}
timeout 600 python bench.py --steps 10 --warmup 3 --frames 96 --template-cache cache/tpl_cfg2.yml.gz > gpurun_out/bench_${R}.log 2>&1; tail -1 gpurun_out/bench_${R}.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python __graft_entry__.py --smoke > gpurun_out/memcheck_${R}.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_${R}.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 3 --frames 96 --no-e2e --no-cpu --template-cache cache/tpl_cfg2.yml.gz > gpurun_out/ncu_launch_${R}.log 2>&1; tail -2 gpurun_out/ncu_launch_${R}.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'similarity_coarse|similarity_local|spread_linearize|cg_quantize' -s 12 -c 8 -o gpurun_out/prof_${R} python bench.py --steps 1 --warmup 3 --frames 96 --no-e2e --no-cpu --template-cache cache/tpl_cfg2.yml.gz > gpurun_out/ncu_full_${R}.log 2>&1; tail -2 gpurun_out/ncu_full_${R}.log
ls -la gpurun_out
