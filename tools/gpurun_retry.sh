#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun flags] -- 'command'   — retries while the pod answers busy (exit code 3), up to 12 times
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
