#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout -s KILL 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu.log
nproc; free -g | head -2
timeout -s KILL 900 python bench.py > gpurun_out/r2_bench_cfg3_n1.json 2> gpurun_out/r2_bench_cfg3_n1.err; tail -3 gpurun_out/r2_bench_cfg3_n1.err; cut -c1-1500 gpurun_out/r2_bench_cfg3_n1.json
timeout -s KILL 900 python bench.py --config 4 > gpurun_out/r2_bench_cfg4_n1.json 2> gpurun_out/r2_bench_cfg4_n1.err; tail -3 gpurun_out/r2_bench_cfg4_n1.err; cut -c1-600 gpurun_out/r2_bench_cfg4_n1.json
timeout -s KILL 1200 python bench.py --config 5 --steps 3 --warmup 3 > gpurun_out/r2_bench_cfg5_n1.json 2> gpurun_out/r2_bench_cfg5_n1.err; tail -3 gpurun_out/r2_bench_cfg5_n1.err; cut -c1-600 gpurun_out/r2_bench_cfg5_n1.json
