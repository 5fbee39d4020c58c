#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout -s KILL 600 python tools/diag_det.py > gpurun_out/r2b_det.log 2>&1
cat gpurun_out/r2b_det.log
SUBRECON_B200_LIB=tools/_build/libsubrecon_tl.so timeout -s KILL 300 python tools/timeline.py > gpurun_out/r2b_tl.log 2>&1
SUBRECON_B200_LIB=tools/_build/libsubrecon_tl.so timeout -s KILL 300 python tools/timeline.py --tables 0 > gpurun_out/r2b_tl_notab.log 2>&1
cat gpurun_out/r2b_tl.log gpurun_out/r2b_tl_notab.log
