#!/bin/bash
# One bench.py JSON line per BASELINE config / arithmetic path (1 GPU).   gpurun --timeout 1200 -- 'bash tools/gpu_configs.sh TAG'
TAG=${1:-c}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 400 python bench.py "$@" 2>> gpurun_out/configs_$TAG.err | grep '^{' > gpurun_out/bench_${name}_$TAG.json
  python - gpurun_out/bench_${name}_$TAG.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print('%-34s %s gemm=%s  %.3f ms  %.4g env-steps/s  e2e %.3f ms  GEMM %.1f TF/s' % (sys.argv[1].split('/')[-1], d['config']['workload'], d['config']['gemm'][:14], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['achieved']))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
run cfg3_tf32x3 --steps 50 --warmup 5
run cfg2_tf32x3 --workload halfcheetah-cfg2 --steps 50 --warmup 5
run cfg3_simt --gemm simt --steps 20 --warmup 3 --no-cpu-baseline
run cfg4_bf16 --workload humanoid-wide-cfg4 --gemm bf16 --steps 5 --warmup 3 --no-cpu-baseline
run cfg4_tf32x3 --workload humanoid-wide-cfg4 --gemm tf32x3 --steps 5 --warmup 3 --no-cpu-baseline
DPPO_TC_TN_WIDE=1 run cfg3_tnwide --steps 30 --warmup 5 --no-cpu-baseline
tail -3 gpurun_out/configs_$TAG.err
