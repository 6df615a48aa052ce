#!/bin/bash
# full ncu capture of N launches matching a regex, starting after SKIP matching launches
TAG=$1; RX=$2; SKIP=$3; N=$4; shift 4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c $N \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > /dev/null 2> gpurun_out/prof_$TAG.err
ls -la gpurun_out/prof_$TAG.ncu-rep; tail -2 gpurun_out/prof_$TAG.err
