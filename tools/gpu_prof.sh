#!/bin/bash
# ncu visit: launch list + full captures of the tensor-core kernels + a clean bench line.   gpurun -- 'bash tools/gpu_prof.sh TAG [bench args]'
TAG=${1:-p}; shift
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 50 -c 34 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > /dev/null 2> gpurun_out/prof_$TAG.err
# one full-set capture of every tensor-core GEMM launch of one step (fwd x3, dX x2, dW x3)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"h_gemm_kernel|h_dw_kernel|tc_gemm_kernel|tc_tn_tma_kernel" -s 24 -c 8 \
    -o gpurun_out/prof_tc_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > /dev/null 2>> gpurun_out/prof_$TAG.err
timeout 300 python bench.py --steps 50 --warmup 5 "$@" > gpurun_out/bench_$TAG.json 2>> gpurun_out/prof_$TAG.err
ls -la gpurun_out | grep $TAG; tail -3 gpurun_out/prof_$TAG.err
