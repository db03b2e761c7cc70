#!/bin/bash
# usage: retry.sh <timeout_s> <script> <log>   - retries a gpurun call while the pod answers "busy" (exit code 3)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" > "$3" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
