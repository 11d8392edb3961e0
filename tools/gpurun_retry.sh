#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...>   -- retries while gpurun answers "busy" (exit code 3)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
