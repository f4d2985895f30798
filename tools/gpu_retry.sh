#!/bin/bash
# tools/gpu_retry.sh <logfile> <gpurun args...>: retry a gpurun call while the pod answers "busy" (exit code 3)
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 100
done
exit 3
