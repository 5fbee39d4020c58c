#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout -s KILL 600 python tools/diag_det.py 2>&1 | grep -v "^$" | cut -c1-220 > gpurun_out/r2d_det.log
cat gpurun_out/r2d_det.log
for args in "--cols 1000000" "--cols 1000000 --k-fastest 0" "--cols 1000000 --tables 0" "--cols 1000000 --tables 0 --k-fastest 0" "--tree caterpillar --taxa 5000 --ncat 8 --cols 200000" "--tree balanced --cols 1000000"; do
  timeout -s KILL 300 python tools/prune_probe.py $args --reps 3 >> gpurun_out/r2d_probe.log 2>&1
done
cat gpurun_out/r2d_probe.log
SUBRECON_B200_LIB=tools/_build/libsubrecon_tl.so timeout -s KILL 300 python tools/timeline.py > gpurun_out/r2d_tl.log 2>&1
SUBRECON_B200_LIB=tools/_build/libsubrecon_tl.so timeout -s KILL 300 python tools/timeline.py --tables 0 > gpurun_out/r2d_tl_notab.log 2>&1
cat gpurun_out/r2d_tl.log gpurun_out/r2d_tl_notab.log
