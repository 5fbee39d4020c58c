#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> [--gpus N] -- <command>   (retries while the pod answers busy)
log=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
  if grep -q "status=transient\|rc=3" $log && ! grep -q "status=ok" $log; then sleep 45; continue; fi
  break
done
