#!/bin/bash
# gpurun with retries for an arbitrary command.  Usage: tools/gpurun_retry_cmd.sh <tag> <timeout_s> <gpus> <command...>
tag=$1; tmo=$2; gpus=$3; shift 3
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" --gpus "$gpus" -- "$@" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/${tag}_call.log || [ $rc -eq 3 ]; then
    echo "attempt $attempt: transient (rc=$rc), retrying in 150 s" >> gpurun_out/${tag}_retry.log
    sleep 150
    continue
  fi
  echo "attempt $attempt: done rc=$rc" >> gpurun_out/${tag}_retry.log
  exit $rc
done
exit 3
