#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   retries while the pod answers "busy" (exit 3)
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
