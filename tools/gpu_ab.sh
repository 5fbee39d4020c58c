#!/bin/bash
# A/B of pruning-kernel builds on the probe workloads: usage  bash tools/gpu_ab.sh <lib> [<lib> ...]   (paths inside the repo)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
rm -f $O/r02_ab.log
for L in "$@"; do
  export SUBRECON_B200_LIB=$PWD/$L
  echo "== $L" >> $O/r02_ab.log
  for v in "--cols 1000000 --distinct" "--cols 1000000 --distinct --tables 0" "--tree caterpillar --taxa 5000 --ncat 8 --cols 200000 --model blosum62 --distinct" "--taxa 2000 --cols 1000000 --model dayhoff --distinct" "--tree balanced --cols 1000000 --distinct"; do
    timeout -s KILL 300 python tools/prune_probe.py $v --reps 3 2>&1 | tail -1 | sed 's/k_fastest=-1 //; s/ Mcol.*TFLOP.s//' >> $O/r02_ab.log
  done
done
cat $O/r02_ab.log
