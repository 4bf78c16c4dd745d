#!/bin/bash
# Keep asking for a GPU box until the call actually runs (status "transient" = no slot, nothing charged).
# usage: tools/gpu_retry.sh <timeout-seconds> <command...>
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$OUT" | tail -n 120
  exit 0
done
echo "gave up: no GPU slot"
