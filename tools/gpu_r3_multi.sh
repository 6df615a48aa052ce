#!/bin/bash
# Round-2 N-GPU visit (one gpurun --gpus N call): weak and strong scaling of the default path, BASELINE configs[4] (minibatch
# 256 / 4096 per GPU, graph replay) and configs[3] (4x1024, N=256) on N GPUs in the fp16-split and the bf16 arithmetic, plus
# the NCCL multi-GPU parity test.   gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_r3_multi.sh N TAG'
N=${1:-8}; TAG=${2:-r3m}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$TAG.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29600
run() {   # name timeout args...
  name=$1; tmo=$2; shift 2
  port=$((port+1))
  NCCL_DEBUG=WARN timeout $tmo $RUN --master-port $port bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  echo "$name rc=$? $(cut -c1-260 gpurun_out/${TAG}_${name}.json)"
}
run weak 300 --steps 30 --warmup 5
run strong 300 --steps 30 --warmup 5 --scaling strong
run mb256 200 --minibatch 256 --steps 5 --warmup 3
#run mb4096 200 --minibatch 4096 --steps 5 --warmup 3
run mb256_strong 200 --minibatch 256 --steps 5 --warmup 3 --scaling strong
#run cfg4_fp16x2 400 --workload humanoid-wide-cfg4 --gemm fp16x2 --steps 8 --warmup 3 --no-cpu-baseline
run cfg4_bf16 400 --workload humanoid-wide-cfg4 --gemm bf16 --steps 8 --warmup 3 --no-cpu-baseline
(timeout 400 python -m pytest tests -m gpu -q -k "multi_gpu" 2>&1 | tail -8) > gpurun_out/${TAG}_tests_multi.log
tail -3 gpurun_out/${TAG}_tests_multi.log
