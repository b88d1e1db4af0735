#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status=transient: nothing is charged).
# usage: tools/gpurun_retry.sh <log file> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if ! grep -q "status=transient" "$LOG"; then exit 0; fi
  sleep 90
done
exit 3
