#!/bin/bash
# Dev-time: gpurun answers 3 when no box / slot is free (nothing charged): retry until it is served.
# usage: tools/dev/gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
