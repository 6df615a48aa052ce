#!/bin/bash
# N-GPU visit: weak-scaling bench, strong-scaling bench, minibatch sweep with the exchange timed (BASELINE configs[2], [4]).
# gpurun --gpus N --timeout 900 -- 'bash tools/gpu_scale.sh N TAG'
N=${1:-8}; TAG=${2:-s}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NCCL_DEBUG=WARN timeout 300 $RUN --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/scale_$TAG.err
NCCL_DEBUG=WARN timeout 300 $RUN --master-port 29522 bench.py --gpus $N --steps 20 --warmup 3 --scaling strong \
    > gpurun_out/bench_strong_n${N}_$TAG.json 2>> gpurun_out/scale_$TAG.err
NCCL_DEBUG=WARN timeout 300 $RUN --master-port 29523 tools/minibatch_sweep.py \
    > gpurun_out/minibatch_sweep_n${N}_$TAG.jsonl 2>> gpurun_out/scale_$TAG.err
for f in gpurun_out/bench_n${N}_$TAG.json gpurun_out/bench_strong_n${N}_$TAG.json; do
python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], 'scaling', d['scaling'], 'value %.4e ms %.3f e2e %.4e (%.3f ms) rows/gpu %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['rows_per_gpu']))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
done
cut -c1-330 gpurun_out/minibatch_sweep_n${N}_$TAG.jsonl; tail -4 gpurun_out/scale_$TAG.err
