#!/bin/bash
# first contact of the rewritten prune_kernel with a GPU: small parity tests first (a hang shows early), then timing probes
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg1_lysozyme" > gpurun_out/r2a_t1.log 2>&1 || { tail -30 gpurun_out/r2a_t1.log; exit 1; }
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg2 or clade_table_shapes or fuzz or polytomy or variants or adversarial or degenerate or leaf_as_root or ragged or cherry_tables_match or device_built" > gpurun_out/r2a_t2.log 2>&1
tail -15 gpurun_out/r2a_t2.log
for args in "--cols 1000000" "--cols 1000000 --k-fastest 0" "--cols 1000000 --tables 0" "--cols 1000000 --uniform" "--tree caterpillar --taxa 5000 --ncat 8 --cols 200000" "--tree balanced --cols 1000000"; do
  timeout -s KILL 300 python tools/prune_probe.py $args --reps 3 >> gpurun_out/r2a_probe.log 2>&1
done
cat gpurun_out/r2a_probe.log
