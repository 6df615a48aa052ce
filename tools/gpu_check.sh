#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list and one full capture of the top kernel.
# Usage (from the authoring container):  gpurun --timeout 1500 -- 'bash tools/gpu_check.sh TAG'
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt
(timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60) > gpurun_out/tests_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
# every launch with its device time (cold cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>> gpurun_out/bench_$TAG.err
# the top kernel, full set, source-level
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sgemm -s 9 -c 8 \
    -o gpurun_out/prof_sgemm_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>> gpurun_out/bench_$TAG.err
tail -5 gpurun_out/tests_$TAG.log; tail -2 gpurun_out/smoke_$TAG.log; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
