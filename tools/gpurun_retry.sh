#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout_s> <command...>   -- retries while the pod answers busy (exit 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 120
done
exit 3
