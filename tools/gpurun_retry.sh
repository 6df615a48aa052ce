#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged).  Usage: tools/gpurun_retry.sh <tag> <timeout_s> [--gpus N]
tag=$1; tmo=$2; shift 2
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" "$@" -- bash tools/gpu_job.sh "$tag" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/${tag}_call.log || [ $rc -eq 3 ]; then
    echo "attempt $attempt: transient (rc=$rc), retrying in 120 s" >> gpurun_out/${tag}_retry.log
    sleep 120
    continue
  fi
  echo "attempt $attempt: done rc=$rc" >> gpurun_out/${tag}_retry.log
  exit $rc
done
exit 3
