#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
python tools/prune_probe.py --cols 262144 --reps 2 > gpurun_out/r2e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:prune_kernel -s 1 -c 1 -o gpurun_out/prof_r02_prune_256k -f python tools/prune_probe.py --cols 262144 --reps 2 > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_plain.log gpurun_out/r2e_ncu.log
