#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> <timeout-seconds> <command string>   -- retries gpurun while the pod answers "transient"/busy
log=$1; shift; to=$1; shift
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpu_retry: rc=$rc attempts=$attempt" >> "$log"
