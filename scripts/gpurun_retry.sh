#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout> [--gpus N] -- '<command>'   (retries while the pod answers busy / transient)
log=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$to" "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$log" || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
echo "gpurun_retry rc=$rc" >> "$log"
