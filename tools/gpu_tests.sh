#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout -s KILL 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1
tail -25 gpurun_out/r2_pytest_gpu.log
