#!/bin/bash
# Multi-GPU visit: bench under torch.distributed.run at N ranks + the 2-rank NCCL parity test.
# gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh N TAG'
N=${1:-2}; TAG=${2:-m}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$TAG.txt
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
(timeout 600 python -m pytest tests -m gpu -q --timeout 500 -k "multi_gpu" 2>&1 | tail -30) > gpurun_out/tests_multi_$TAG.log
tail -5 gpurun_out/tests_multi_$TAG.log
cut -c1-400 gpurun_out/bench_n${N}_$TAG.json; tail -5 gpurun_out/bench_n${N}_$TAG.err
