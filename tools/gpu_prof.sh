#!/bin/bash
# ncu visit: launch list + one full capture of the kernels matching REGEX.   gpurun -- 'bash tools/gpu_prof.sh TAG REGEX [bench args]'
TAG=${1:-p}; RX=${2:-gemm}; shift 2
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > /dev/null 2> gpurun_out/prof_$TAG.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RX -s 14 -c 7 \
    -o gpurun_out/prof_${RX}_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > /dev/null 2>> gpurun_out/prof_$TAG.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_$TAG.json 2>> gpurun_out/prof_$TAG.err
ls -la gpurun_out | grep $TAG; tail -3 gpurun_out/prof_$TAG.err
