#!/bin/sh
# usage: tools/gpurun_retry.sh <log> <gpurun args...>  -- retries while the pod answers "busy" (exit 3, nothing charged)
LOG=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 120
done
exit 3
