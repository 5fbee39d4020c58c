#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; tail -4 $O/r02_pytest_gpu.log
timeout -s KILL 600 python bench.py > $O/r02_bench_cfg3_n1.json 2> $O/r02_bench_cfg3_n1.err; tail -3 $O/r02_bench_cfg3_n1.err; cut -c1-400 $O/r02_bench_cfg3_n1.json
python -c "
import json;d=json.loads(open('$O/r02_bench_cfg3_n1.json').read());print(d['kernel_ms'], d['e2e']['value'])"
timeout -s KILL 600 python tools/parity_report.py --out $O/r02_parity_report.json
for c in 1 2; do timeout -s KILL 600 python bench.py --config $c --steps 20 --warmup 5 > $O/r02_bench_cfg${c}_n1.json 2> $O/r02_bench_cfg${c}_n1.err; tail -3 $O/r02_bench_cfg${c}_n1.err; cut -c1-300 $O/r02_bench_cfg${c}_n1.json; done
