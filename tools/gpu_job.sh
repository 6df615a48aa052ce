#!/bin/bash
# one gpurun call: the job steps are given as "name|timeout|command" lines in tools/gpu_job.txt; each step's output goes to
# gpurun_out/<tag>_<name>.log.  Usage: gpurun -- bash tools/gpu_job.sh <tag>
tag=${1:-job}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.log 2>&1
while IFS='|' read -r name tmo cmd; do
  [ -z "$name" ] && continue
  case "$name" in \#*) continue;; esac
  echo "=== $name (timeout $tmo s): $cmd"
  start=$(date +%s)
  timeout "$tmo" bash -c "$cmd" > gpurun_out/${tag}_${name}.log 2>&1
  rc=$?
  echo "rc=$rc  ($(( $(date +%s) - start )) s)" | tee -a gpurun_out/${tag}_${name}.log
  tail -n 4 gpurun_out/${tag}_${name}.log
done < tools/gpu_job.txt
