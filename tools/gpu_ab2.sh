#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
bash tools/gpu_ab.sh "$@"
M=gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
for L in "$@"; do
  export SUBRECON_B200_LIB=$PWD/$L
  echo "== ncu $L"
  timeout -s KILL 600 ncu --metrics $M --clock-control none -k regex:prune_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_ab_ncu.csv python tools/prune_probe.py --cols 1000000 --distinct --reps 2 > /dev/null 2>&1
  grep -v "^==" gpurun_out/r02_ab_ncu.csv | cut -d, -f13- | tail -6 | tr '\n' ' '; echo
done
