#!/bin/bash
# tests + the default bench line on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; tail -4 $O/r02_pytest_gpu.log
timeout -s KILL 600 python bench.py > $O/r02_bench_cfg3_n1.json 2> $O/r02_bench_cfg3_n1.err; tail -3 $O/r02_bench_cfg3_n1.err; cut -c1-1200 $O/r02_bench_cfg3_n1.json
