#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
run() {  # run <n> <tag> <bench args...>
  n=$1; tag=$2; shift 2
  SECONDS=0
  timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --no-cpu-baseline "$@" > $O/r02_bench_$tag.json 2> $O/r02_bench_$tag.err
  echo "== $tag rc=$? wall $SECONDS s"; cut -c1-200 $O/r02_bench_$tag.json; grep -v "^\[W\|^W[0-9]\|^$\|NCCL version\|^\*\*\*\|OMP_NUM" $O/r02_bench_$tag.err | tail -3
}
run 8 cfg3_n8_weak --steps 5 --warmup 3
run 8 cfg3_n8_strong --scaling strong --steps 10 --warmup 3
run 4 cfg3_n4_weak --steps 5 --warmup 3
run 2 cfg3_n2_weak --steps 5 --warmup 3
