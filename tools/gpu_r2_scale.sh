#!/bin/bash
# Round-2 multi-GPU measurements on ONE 8-GPU box: config #3 weak at N=8 (per-process, single-process and full-joint e2e legs),
# config #5 (2,000 taxa x 8M sites) strong at N=2,4,8, config #3 strong at N=8.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
nvidia-smi -L | head -8; nproc; free -g | head -2
run() {  # run <n> <tag> <bench args...>
  n=$1; tag=$2; shift 2
  timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --no-cpu-baseline "$@" > $O/r02_bench_$tag.json 2> $O/r02_bench_$tag.err
  echo "== $tag rc=$?"; cut -c1-200 $O/r02_bench_$tag.json; grep -v "^\[W\|^W[0-9]\|^$\|NCCL version" $O/r02_bench_$tag.err | tail -3
}
run 8 cfg3_n8_weak --steps 5 --warmup 3
run 8 cfg5_n8_strong --config 5 --scaling strong --steps 5 --warmup 3
run 4 cfg5_n4_strong --config 5 --scaling strong --steps 4 --warmup 3
run 2 cfg5_n2_strong --config 5 --scaling strong --steps 3 --warmup 3
run 8 cfg3_n8_strong --scaling strong --steps 10 --warmup 3
