#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <gpurun args...>   -- retries while the pod answers "transient" (nothing charged)
LOG=$1; shift
for attempt in $(seq 1 40); do
  gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG" || grep -q "exit code 3" "$LOG"; then sleep 45; continue; fi
  break
done
echo "attempts: $attempt" >> "$LOG"
