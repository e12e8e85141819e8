#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   - retries while gpurun answers "busy" (exit 3), up to ~40 min
log=$1; shift
for i in $(seq 1 16); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
