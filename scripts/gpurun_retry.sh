#!/bin/bash
# gpurun with retries while the pod answers "transient / busy" (exit code 3 or status=transient)
# usage: scripts/gpurun_retry.sh <timeout> <command...>
T=$1; shift
for i in $(seq 1 30); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|nothing was charged"; then sleep 45; continue; fi
  echo "$OUT"
  exit 0
done
echo "$OUT"
