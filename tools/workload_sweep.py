#!/usr/bin/env python
"""BASELINE.json configs beyond the headline one, and the minibatch sweep (configs[4]), on one GPU.
   python tools/workload_sweep.py [--gemm tf32x3|simt]       (run on the GPU box; prints one JSON line per case)"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tensorflow-dppo_b200"))
import torch
from dppo_b200 import DppoEngine
from dppo_b200.synthetic import WORKLOADS, rollout_buffer, params_pair

ap = argparse.ArgumentParser(); ap.add_argument("--gemm", default="tf32x3"); ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
mode = {"simt": 0, "tf32x3": 1, "bf16": 2}[args.gemm]


def time_fn(fn, k):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


for name in ("pendulum-cfg1", "halfcheetah-cfg2", "humanoid-cfg3", "humanoid-wide-cfg4"):
    T, N, obs, act, hidden = WORKLOADS[name]
    B = T * N
    eng = DppoEngine(obs, act, hidden, B, gemm_mode=mode)
    buf = {k: v.cuda() for k, v in rollout_buffer(T, N, obs, act).items()}
    p_pi, p_old = (t.cuda() for t in params_pair(eng.layout))
    m = torch.zeros_like(p_pi); v = torch.zeros_like(p_pi); g = eng.new_grad_buffer()
    st = {"t": 0}

    def step():
        st["t"] += 1
        eng.update(p_pi, p_old, buf["obs"], buf["act"], buf["rew"], buf["predv"], buf["done"], 0.99, 0.95, 0.7,
                   grad_out=g, adam_m=m, adam_v=v, step=st["t"], lr=2e-5 * 0.7)
    ms = time_fn(step, args.steps if B < 300000 else 5)
    print(json.dumps({"case": name, "gemm": args.gemm, "rows": B, "n_params": eng.n_params, "ms_per_update": ms,
                      "env_steps_per_s": B / ms * 1e3}))
    if name == "humanoid-cfg3":   # configs[4]: contiguous minibatch slices 256 * 2^k of the same buffer
        adv, ret = eng.gae(buf["rew"], buf["predv"], buf["done"], 0.99, 0.95)
        adv = adv.view(-1); ret = ret.view(-1)
        for mb in (256, 1024, 4096, 16384, 65536):
            def mb_step():
                eng.loss_grad(p_pi, p_old, buf["obs"][:mb], buf["act"][:mb], adv[:mb], ret[:mb], 0.7, grad_out=g)
                eng.adam_apply(p_pi, g, m, v, 1, 1e-5)
            ms = time_fn(mb_step, 20)
            print(json.dumps({"case": f"humanoid-cfg3 minibatch {mb}", "gemm": args.gemm, "rows": mb, "ms_per_update": ms,
                              "env_steps_per_s": mb / ms * 1e3, "alg_GBps": mb * 1600 / ms / 1e6}))
    del eng
    torch.cuda.empty_cache()
