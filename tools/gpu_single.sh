#!/bin/bash
# ONE host process driving all 8 GPUs (sr_multi_*), no other process on the GPUs: the route a single JVM takes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout -s KILL 600 python bench.py --single-process-devices 8 --no-cpu-baseline --no-full-joint --steps 5 --warmup 3 > $O/r02_bench_cfg3_single_process_8.json 2> $O/r02_bench_cfg3_single_process_8.err
tail -3 $O/r02_bench_cfg3_single_process_8.err
python -c "
import json;d=json.loads(open('$O/r02_bench_cfg3_single_process_8.json').read());print(d['value'], d['e2e']['value'], d['e2e_single_process'])"
