#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
rm -f gpurun_out/r2f_probe.log
for args in "--cols 1000000" "--cols 1000000 --tables 0" "--tree caterpillar --taxa 5000 --ncat 8 --cols 200000" "--tree balanced --cols 1000000" "--taxa 2000 --cols 1000000"; do
  timeout -s KILL 300 python tools/prune_probe.py $args --reps 3 >> gpurun_out/r2f_probe.log 2>&1
done
cat gpurun_out/r2f_probe.log
timeout -s KILL 600 python tools/diag_det.py 2>&1 | grep -v "^$" | cut -c1-200 > gpurun_out/r2f_det.log; cat gpurun_out/r2f_det.log
M=sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed
for t in 1 0; do
python tools/prune_probe.py --cols 262144 --reps 2 --tables $t > gpurun_out/r2f_plain.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:prune_kernel -s 1 -c 1 --csv --log-file gpurun_out/r2f_ncu_t$t.csv python tools/prune_probe.py --cols 262144 --reps 2 --tables $t > /dev/null 2>&1
grep -v "^==" gpurun_out/r2f_ncu_t$t.csv | cut -d, -f5,13- 
done
