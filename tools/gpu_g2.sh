#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; tail -4 $O/r02_pytest_gpu.log
for c in 3 4 5; do
  /usr/bin/time -f "cfg$c wall %e s" timeout -s KILL 900 python bench.py --config $c > $O/r02_bench_cfg${c}_n1.json 2> $O/r02_bench_cfg${c}_n1.err; tail -2 $O/r02_bench_cfg${c}_n1.err
  python -c "
import json;d=json.loads(open('$O/r02_bench_cfg${c}_n1.json').read());print(d['value'], d['e2e']['value'], d['kernel_ms'], d['roofline']['frac'], d['roofline']['frac_executed'])"
done
