#!/bin/bash
# usage: tools/gpurun_retry.sh <max_tries> <gpurun args...>   -- retries while gpurun answers "busy" (exit 3, nothing charged)
tries=$1; shift
for i in $(seq 1 $tries); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy (attempt $i), sleeping"; sleep 150
done
exit 3
