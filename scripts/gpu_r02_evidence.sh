#!/bin/bash
# round 2 evidence session (1 GPU): pytest -m gpu, smoke, bench (both arms), ncu launch list of the bench command,
# ncu --set full of every kernel family, compute-sanitizer memcheck on smoke + the frame-side parity tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LMB200_QUIET=1
R=${1:-r02}
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 | tee gpurun_out/${R}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/${R}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 --template-cache gpurun_out/${R}_tpl_cache.yml.gz > gpurun_out/${R}_bench.log 2>&1; tail -1 gpurun_out/${R}_bench.log > gpurun_out/${R}_bench_line.json; cut -c1-400 gpurun_out/${R}_bench_line.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_reference.log 2>&1; tail -1 gpurun_out/${R}_bench_reference.log > gpurun_out/${R}_bench_reference_line.json; cut -c1-300 gpurun_out/${R}_bench_reference_line.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-ts --no-extra --template-cache gpurun_out/${R}_tpl_cache.yml.gz > gpurun_out/${R}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'similarity_coarse|similarity_local|spread_strip|spread_flat|cg_quantize2|dn_median|pyrdown_planar|pack_kernel' -s 33 -c 22 -o gpurun_out/${R}_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-ts --no-extra --template-cache gpurun_out/${R}_tpl_cache.yml.gz > gpurun_out/${R}_ncu_full.log 2>&1; tail -1 gpurun_out/${R}_ncu_full.log | cut -c1-200
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python __graft_entry__.py --smoke 2>&1 | tail -4 | tee gpurun_out/${R}_memcheck_smoke.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 -k "device_epilogue or frame_side_synthetic or masks" 2>&1 | tail -4 | tee gpurun_out/${R}_memcheck_tests.log
