#!/bin/bash
# A/B: the round-2 kernel as committed before gather-kind specialisation (tools/_build/libsr_shipped.so) against the tree's build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
rm -f $O/r02_spec.log
for L in tools/_build/libsr_shipped.so subrecon_b200/libsubrecon_b200.so; do
  export SUBRECON_B200_LIB=$PWD/$L
  echo "== $L" >> $O/r02_spec.log
  for v in "--cols 1000000 --distinct" "--cols 1000000 --distinct --tables 0" "--tree caterpillar --taxa 5000 --ncat 8 --cols 200000 --model blosum62 --distinct" "--taxa 2000 --cols 1000000 --model dayhoff --distinct" "--tree balanced --cols 1000000 --distinct"; do
    timeout -s KILL 300 python tools/prune_probe.py $v --reps 3 2>&1 | tail -1 >> $O/r02_spec.log
  done
done
timeout -s KILL 600 python tools/quick_parity.py >> $O/r02_spec.log 2>&1
M=sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
timeout -s KILL 600 ncu --metrics $M --clock-control none -k regex:prune_kernel -s 1 -c 1 --csv --log-file $O/r02_spec2_ncu.csv python tools/prune_probe.py --cols 1000000 --distinct --reps 2 > /dev/null 2>&1
grep -v "^==" $O/r02_spec2_ncu.csv | cut -d, -f13- | tail -4 >> $O/r02_spec.log
cat $O/r02_spec.log
