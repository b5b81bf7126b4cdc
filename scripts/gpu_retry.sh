#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout_s> <script> <log> [gpus]  -- retries while the pod answers busy/transient
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun ${4:+--gpus $4} --timeout $1 -- "bash $2" > $3 2>&1
  if grep -qE "status=(transient|busy)|exit code 3|rc=3" $3; then sleep 90; continue; fi
  break
done
