#!/bin/bash
# Round-2 evidence refresh on ONE GPU: tests, the bench line + its ncu launch list, a --set full capture of the shipped
# prune_kernel at 1M columns, the same kernel on 1M DISTINCT simulated columns, cuBLAS DGEMM next to the DMMA probe.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; tail -4 $O/r02_pytest_gpu.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout -s KILL 600 $B > $O/r02_bench_line_same_command.json 2> $O/r02_bench_line_same_command.err && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv $B > $O/r02_bench_under_ncu.log 2>&1
cut -c1-300 $O/r02_bench_line_same_command.json
P="python tools/prune_probe.py --cols 1000000 --reps 2"
timeout -s KILL 300 $P > $O/r02_probe_1m.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:prune_kernel -s 1 -c 1 -o $O/prof_r02_prune_1m -f $P > $O/r02_ncu_full_1m.log 2>&1
tail -2 $O/r02_probe_1m.log $O/r02_ncu_full_1m.log
M=gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
for v in "" "--distinct"; do
  tag=tiled; [ -n "$v" ] && tag=distinct
  timeout -s KILL 600 $P $v > $O/r02_probe_1m_$tag.log 2>&1 && \
  timeout -s KILL 900 ncu --metrics $M --clock-control none -k regex:prune_kernel -s 1 -c 1 --csv --log-file $O/r02_ncu_1m_$tag.csv $P $v > /dev/null 2>&1
  cat $O/r02_probe_1m_$tag.log; grep -v "^==" $O/r02_ncu_1m_$tag.csv | cut -d, -f13- | tail -9
done
timeout -s KILL 900 python bench.py --distinct --no-cpu-baseline > $O/r02_bench_cfg3_n1_distinct.json 2> $O/r02_bench_cfg3_n1_distinct.err; cut -c1-400 $O/r02_bench_cfg3_n1_distinct.json
timeout -s KILL 300 python tools/dgemm_peak.py --out $O/fp64_peak_r02.json 2>&1 | tail -1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv,noheader
