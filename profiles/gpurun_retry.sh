#!/bin/bash
# usage: gpurun_retry.sh <timeout-seconds> <command...>   -- retries while the pod answers "busy" (exit code 3)
T="$1"; shift
for try in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
