#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3).  Usage: tools/gpurun_retry.sh <timeout_s> '<command>'
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null || true
  sleep 90
done
exit 3
