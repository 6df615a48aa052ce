"""Frozen synthetic rollout buffers and parameters for benches and smoke runs.

There is no simulator (gym) in this image; the benchmark feeds buffers of the shape the
reference's Worker produces (Worker.py:52-56,94-95): obs[T*N,obs], act[T*N,A], rew/predv/done[T,N].
"""
import numpy as np
import torch

from .layout import ParamLayout

WORKLOADS = {
    # name: (T, N, obs_dim, act_dim, hidden)        BASELINE.json configs[0..3]
    "pendulum-cfg1": (200, 4, 3, 1, (16,)),
    "halfcheetah-cfg2": (2048, 32, 17, 6, (256, 256)),
    "humanoid-cfg3": (2048, 64, 376, 17, (256, 256)),
    "humanoid-wide-cfg4": (2048, 256, 376, 17, (1024, 1024, 1024, 1024)),
}


def rollout_buffer(T, N, obs_dim, act_dim, seed=1234, pin=False):
    """Host tensors, fp32.  obs/act/rew ~ N(0,1); done ~ Bernoulli(1/200); predv ~ N(0,1)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    B = T * N
    buf = dict(
        obs=torch.randn(B, obs_dim, generator=g, dtype=torch.float32),
        act=torch.randn(B, act_dim, generator=g, dtype=torch.float32),
        rew=torch.randn(T, N, generator=g, dtype=torch.float32),
        done=(torch.rand(T, N, generator=g) < (1.0 / 200.0)).to(torch.float32),
        predv=torch.randn(T, N, generator=g, dtype=torch.float32),
    )
    if pin:
        buf = {k: v.pin_memory() for k, v in buf.items()}
    return buf


def normc(shape, std, rng):
    """Column-normalised Gaussian initialiser (the formula of tf_util.py:286-291)."""
    w = rng.standard_normal(shape).astype(np.float32)
    w *= std / np.sqrt(np.square(w).sum(axis=0, keepdims=True))
    return w


def params_pair(layout: ParamLayout, trunk_std=1.0, head_std=0.1, delta=0.02, seed_old=7, seed_pi=8):
    """theta_old: normc kernels + zero biases; theta_pi = theta_old + delta*N(0,1)  (pi != oldpi so that
    both clip branches, both max branches and the ReLU mask are exercised)."""
    rng = np.random.default_rng(seed_old)
    old = np.zeros(layout.n_params, np.float32)
    for e in layout.entries:
        if e.name[0] == "W":
            std = head_std if e.name in ("Wv", "Wp") else trunk_std
            old[e.offset:e.offset + e.size] = normc(e.shape, std, rng).ravel()
    noise = np.random.default_rng(seed_pi).standard_normal(layout.n_params).astype(np.float32)
    pi = (old + delta * noise).astype(np.float32)
    return torch.from_numpy(pi), torch.from_numpy(old)
