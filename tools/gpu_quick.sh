#!/bin/bash
# Quick GPU-box visit: parity tests + bench (no profiler).   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh TAG [pytest -k expr]'
TAG=${1:-q}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 500 ${2:+-k "$2"} 2>&1 | tail -40) > gpurun_out/tests_$TAG.log
timeout 400 python bench.py --steps 30 --warmup 5 ${GEMM:+--gemm $GEMM} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
tail -8 gpurun_out/tests_$TAG.log
cat gpurun_out/bench_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.3e  ms/step %.3f  e2e %.3e (%.3f ms)  gemm TF/s %.1f  launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['gpu_launches']))
print('families', {k: round(v,3) for k,v in d['roofline']['family_ms_per_step'].items()})
print('cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'), 'clocks', d['clocks'])
"
cut -c1-200 gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
