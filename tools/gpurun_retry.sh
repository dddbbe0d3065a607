#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status=transient).
#   tools/gpurun_retry.sh <log file> <gpurun args...>
LOG=$1; shift
for attempt in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient" "$LOG" || [ $rc -eq 3 ]; then
    sleep 120
    continue
  fi
  exit $rc
done
exit 3
