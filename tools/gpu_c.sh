#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for v in base late fence noprobe base; do
  echo "=== $v"; SUBRECON_B200_LIB=tools/_build/lib_$v.so timeout -s KILL 300 python tools/diag_det.py 2>&1 | grep -v "^$" | cut -c1-220
done > gpurun_out/r2c_det.log 2>&1
cat gpurun_out/r2c_det.log
