#!/bin/bash
# tools/gpurun_retry.sh <timeout-seconds> <command...>: retries while the pod answers "busy" (exit 3), at most 12 times
T=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
