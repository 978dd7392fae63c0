#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <out file> <timeout> <command...>: retries while the pod answers busy (rc 3 / transient)
out=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$to" -- "$@" > "$out" 2>&1
  if grep -q "status=transient\|status=busy\|no box" "$out"; then sleep 90; continue; fi
  break
done
