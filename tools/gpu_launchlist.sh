#!/bin/bash
# ncu launch list of the library's kernels inside the bench command (the data generator's ~10^5 torch launches are filtered out)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout -s KILL 600 $B > $O/r02_bench_line_same_command.json 2> $O/r02_bench_line_same_command.err && \
timeout -s KILL 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:prune_kernel|root_kernel|cherry_kernel|pbuild_kernel|clade_index_kernel|sanitize_kernel|encode_chars_kernel|dmma_peak_kernel" -c 600 --csv --log-file $O/r02_launches_bench.csv $B > $O/r02_bench_under_ncu.log 2>&1
cut -c1-200 $O/r02_bench_line_same_command.json; grep -c prune_kernel $O/r02_launches_bench.csv
