#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> [--gpus N] -- '<command>'   : gpurun with retries while the pod has no free slot
T=$1; shift
for i in $(seq 1 15); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > /tmp/gpu_retry_last.txt 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpu_retry_last.txt || [ $rc -eq 3 ]; then
    echo "[gpu_retry] attempt $i: no slot, sleeping 100 s" >&2
    sleep 100
    continue
  fi
  cat /tmp/gpu_retry_last.txt
  exit $rc
done
cat /tmp/gpu_retry_last.txt
exit 3
