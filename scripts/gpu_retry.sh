#!/bin/bash
# Retries one gpurun call until the pod has a slot (exit 3 / transient = nothing charged).
# usage: scripts/gpu_retry.sh <log> <timeout-s> <command...>
log=$1; to=$2; shift; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient\|nothing was charged" $log; then exit 0; fi
  sleep 200
done
