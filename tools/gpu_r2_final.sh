#!/bin/bash
# Final round-2 evidence on ONE GPU, all from the shipped library: tests, the bench line + its ncu launch list, a --set full
# capture of prune_kernel at 1M distinct columns, metric passes for tiled columns / no tables / the config-#4 caterpillar,
# cuBLAS DGEMM next to the DMMA probe, and the full default bench line (with the CPU baseline).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; tail -3 $O/r02_pytest_gpu.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout -s KILL 600 $B > $O/r02_bench_line_same_command.json 2> $O/r02_bench_line_same_command.err && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:prune_kernel|root_kernel|cherry_kernel|pbuild_kernel|clade_index_kernel|sanitize_kernel|encode_chars_kernel|dmma_peak_kernel" -c 600 --csv --log-file $O/r02_launches_bench.csv $B > $O/r02_bench_under_ncu.log 2>&1
cut -c1-200 $O/r02_bench_line_same_command.json
P="python tools/prune_probe.py --cols 1000000 --reps 2"
timeout -s KILL 300 $P --distinct > $O/r02_probe_1m_distinct.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:prune_kernel -s 1 -c 1 -o $O/prof_r02_prune_1m_distinct -f $P --distinct > $O/r02_ncu_full_1m.log 2>&1
tail -1 $O/r02_probe_1m_distinct.log; tail -2 $O/r02_ncu_full_1m.log
M=gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
i=0
for v in "--cols 1000000" "--cols 1000000 --distinct" "--cols 1000000 --distinct --tables 0" "--tree caterpillar --taxa 5000 --ncat 8 --cols 200000 --model blosum62 --distinct" "--taxa 2000 --cols 1000000 --model dayhoff --distinct"; do
  i=$((i+1))
  timeout -s KILL 600 python tools/prune_probe.py $v --reps 2 > $O/r02_probe_v$i.log 2>&1 && \
  timeout -s KILL 900 ncu --metrics $M --clock-control none -k regex:prune_kernel -s 1 -c 1 --csv --log-file $O/r02_ncu_v$i.csv python tools/prune_probe.py $v --reps 2 > /dev/null 2>&1
  tail -1 $O/r02_probe_v$i.log; grep -v "^==" $O/r02_ncu_v$i.csv | cut -d, -f13- | tail -8 | tr '\n' ' '; echo
done
timeout -s KILL 300 python tools/dgemm_peak.py --out $O/fp64_peak_r02.json 2>&1 | tail -1 | cut -c1-300
timeout -s KILL 900 python bench.py > $O/r02_bench_cfg3_n1.json 2> $O/r02_bench_cfg3_n1.err; tail -2 $O/r02_bench_cfg3_n1.err; cut -c1-300 $O/r02_bench_cfg3_n1.json
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv,noheader
