#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
for c in 4 5; do
  timeout -s KILL 900 python bench.py --config $c > $O/r02_bench_cfg${c}_n1.json 2> $O/r02_bench_cfg${c}_n1.err; tail -2 $O/r02_bench_cfg${c}_n1.err
  python -c "
import json;d=json.loads(open('$O/r02_bench_cfg${c}_n1.json').read());print(d['value'], d['e2e']['value'], d['kernel_ms'], d['roofline']['frac'], d['roofline']['frac_executed'])"
done
