#!/usr/bin/env python
"""A/B of the host-buffer path on one GPU: Adam inside dppo_update_host vs a separate dppo_adam_apply call after it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tensorflow-dppo_b200"))
import torch
from dppo_b200 import DppoEngine
from dppo_b200.synthetic import WORKLOADS, rollout_buffer, params_pair

T, N, obs, act, hidden = WORKLOADS["humanoid-cfg3"]
eng = DppoEngine(obs, act, hidden, T * N, gemm_mode=1)
host = rollout_buffer(T, N, obs, act, seed=1234, pin=True)
p_pi, p_old = (t.cuda() for t in params_pair(eng.layout))
m = torch.zeros_like(p_pi); v = torch.zeros_like(p_pi); g = eng.new_grad_buffer()
st = {"t": 0}


def inside():
    st["t"] += 1
    eng.update_host(p_pi, p_old, host["obs"], host["act"], host["rew"], host["predv"], host["done"], 0.99, 0.95, 0.7,
                    grad_out=g, adam_m=m, adam_v=v, step=st["t"], lr=1e-6)


def outside():
    st["t"] += 1
    eng.update_host(p_pi, p_old, host["obs"], host["act"], host["rew"], host["predv"], host["done"], 0.99, 0.95, 0.7, grad_out=g)
    eng.adam_apply(p_pi, g, m, v, st["t"], 1e-6)


def timed(fn, k=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


for rnd in range(3):
    print(f"round {rnd}: adam inside {timed(inside):.3f} ms   adam outside {timed(outside):.3f} ms", flush=True)
