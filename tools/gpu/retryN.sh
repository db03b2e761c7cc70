#!/bin/bash
# usage: retryN.sh <ngpus> <timeout_s> <script> <log>
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus "$1" --timeout "$2" -- "bash $3" > "$4" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
