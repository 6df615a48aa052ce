"""CPU oracle for the DPPO update path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product path
(``tensorflow-dppo_b200/``) never does: it fails loudly if the CUDA library
is missing.

This is a numpy restatement of the reference's algorithm for the hot path
(the reference itself needs TensorFlow 1.x + gym, neither installed here):

  * actor-critic MLP forward            /root/reference/Model.py:7-18
  * normc initialiser                   /root/reference/Others/tf_util.py:286-291
  * DiagGaussianPd neglogp/kl/entropy   /root/reference/Others/distributions.py:184-208
  * PPO clipped loss                    /root/reference/PPO.py:19,29-40
  * gradient of the loss                /root/reference/PPO.py:46  (TF autodiff; closed form here)
  * tower gradient mean                 /root/reference/PPO.py:55-65
  * Adam apply (TF 1.x form)            /root/reference/PPO.py:20,53
  * GAE + returns + standardisation     /root/reference/Worker.py:82-92
  * lr / clip annealing                 /root/reference/Worker.py:77-80, PPO.py:18-20

Pinning status (see DESIGN.md "Oracle"):
  * GAE: pinned -- ``oracle/make_golden.py`` executes Worker.py:82-92 verbatim
    (numpy only) and the result is committed in tests/golden/gae_ref.npz.
  * DiagGaussianPd, the loss forward values, Model.FC forward and
    normc_initializer: pinned on the reference's own *composition* -- the
    reference source lines are executed over an eager numpy stand-in for the
    ``tf`` ops they call (oracle/tf_numpy_shim.py); TensorFlow's kernels
    themselves are absent, so their op semantics are the restated part.
  * gradients, tower mean, Adam: PARITY UNPINNED by the reference (it has no
    test or golden for them and TF autodiff cannot run here).  The closed form
    below is cross-checked against torch autograd of the same forward in
    tests/test_oracle.py.

Conventions (canonical mode): rank-2 tensors, row = t*N+n,
weights ``[in,out]`` row-major, flat parameter order
``[W1,b1,...,WL,bL,Wv,bv,Wp,bp]`` (Model.py:17 creation order, trunk
generalised to L hidden layers).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------
# parameter layout
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class Layout:
    """Flat parameter layout, Model.py:12-14,17 order (trunk, value head, policy head)."""
    obs_dim: int
    hidden: Tuple[int, ...]
    act_dim: int
    entries: Tuple[Tuple[str, Tuple[int, ...], int], ...] = field(default=())
    n_params: int = 0
    pd: str = "gauss"          # "gauss": DiagGaussianPd, act_dim = A, policy head 2A wide (distributions.py:84-85)
                               # "cat":   CategoricalPd,  act_dim = ncat, policy head ncat wide (distributions.py:53-54)

    @property
    def pd_width(self) -> int:
        return 2 * self.act_dim if self.pd == "gauss" else self.act_dim


def make_layout(obs_dim: int, hidden: Sequence[int], act_dim: int, pd: str = "gauss") -> Layout:
    entries = []
    off = 0
    fan_in = obs_dim
    for l, h in enumerate(hidden):
        entries.append((f"W{l + 1}", (fan_in, h), off)); off += fan_in * h
        entries.append((f"b{l + 1}", (h,), off)); off += h
        fan_in = h
    entries.append(("Wv", (fan_in, 1), off)); off += fan_in
    entries.append(("bv", (1,), off)); off += 1
    assert pd in ("gauss", "cat")
    pw = 2 * act_dim if pd == "gauss" else act_dim                # Model.py:14 pdtype.param_shape()[0]
    entries.append(("Wp", (fan_in, pw), off)); off += fan_in * pw
    entries.append(("bp", (pw,), off)); off += pw
    return Layout(obs_dim, tuple(hidden), act_dim, tuple(entries), off, pd)


def unflatten(theta: np.ndarray, lay: Layout) -> dict:
    out = {}
    for name, shape, off in lay.entries:
        n = int(np.prod(shape))
        out[name] = theta[off:off + n].reshape(shape)
    return out


def normc_columns(shape, std, rng) -> np.ndarray:
    """tf_util.py:286-291: Gaussian, every column L2-normalised to ``std``."""
    out = rng.standard_normal(shape).astype(np.float32)
    out *= std / np.sqrt(np.square(out).sum(axis=0, keepdims=True))
    return out


def init_params(lay: Layout, rng, trunk_std=0.01, head_std=0.01) -> np.ndarray:
    """Model.py:10 (normc(0.01) kernels, zero biases); std overridable for test buffers."""
    theta = np.zeros(lay.n_params, np.float32)
    for name, shape, off in lay.entries:
        if name[0] == "W":
            std = head_std if name in ("Wv", "Wp") else trunk_std
            n = int(np.prod(shape))
            theta[off:off + n] = normc_columns(shape, std, rng).ravel()
    return theta


# --------------------------------------------------------------------------
# MLP forward   (Model.py:11-15)
# --------------------------------------------------------------------------
def bf16_round(x):
    """Round-to-nearest-even to bfloat16, returned in the input's float dtype (models the operand rounding of the
    DPPO_GEMM_BF16 path: every matmul operand is rounded, products are exact, accumulation is wide)."""
    import torch
    a = np.asarray(x)
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16).to(torch.float32).numpy().astype(a.dtype)


def _ident(x):
    return x


def mlp_forward(theta, obs, lay: Layout, dtype=np.float64, q=_ident):
    """Returns v[B], flat[B,2A], list of post-ReLU hidden activations.  q = operand rounding of every matmul."""
    p = unflatten(np.asarray(theta, dtype), lay)
    h = np.asarray(obs, dtype)
    hs = []
    for l in range(len(lay.hidden)):
        h = np.maximum(q(h) @ q(p[f"W{l + 1}"]) + p[f"b{l + 1}"], 0)   # Model.py:12 dense(relu)
        hs.append(h)
    v = (q(h) @ q(p["Wv"]) + p["bv"])[:, 0]                       # Model.py:13
    flat = q(h) @ q(p["Wp"]) + p["bp"]                            # Model.py:14
    return v, flat, hs


# --------------------------------------------------------------------------
# DiagGaussianPd  (distributions.py:184-208)
# --------------------------------------------------------------------------
def dg_split(flat):
    A = flat.shape[-1] // 2
    return flat[..., :A], flat[..., A:]                           # :187


def dg_neglogp(flat, x):
    mean, logstd = dg_split(flat)
    std = np.exp(logstd)                                          # :190
    A = x.shape[-1]
    return (0.5 * np.sum(np.square((x - mean) / std), axis=-1)
            + 0.5 * LOG_2PI * A + np.sum(logstd, axis=-1))        # :195-198


def dg_logp(flat, x):
    return -dg_neglogp(flat, x)                                   # :25-26


def dg_entropy(flat):
    _, logstd = dg_split(flat)
    return np.sum(logstd + 0.5 * math.log(2.0 * math.pi * math.e), axis=-1)   # :202-203


def dg_kl(flat_p, flat_q):
    mp, lp = dg_split(flat_p)
    mq, lq = dg_split(flat_q)
    return np.sum(lq - lp + (np.square(np.exp(lp)) + np.square(mp - mq))
                  / (2.0 * np.square(np.exp(lq))) - 0.5, axis=-1)  # :199-201


def dg_sample(flat, eps):
    mean, logstd = dg_split(flat)
    return mean + np.exp(logstd) * eps                            # :204-205


def dg_mode(flat):
    return dg_split(flat)[0]                                      # :193-194


# --------------------------------------------------------------------------
# CategoricalPd  (distributions.py:124-159)
# --------------------------------------------------------------------------
def _cat_parts(logits):
    a0 = logits - np.max(logits, axis=-1, keepdims=True)          # :141,149
    ea0 = np.exp(a0)
    z0 = np.sum(ea0, axis=-1, keepdims=True)
    return a0, ea0, z0


def cat_neglogp(logits, x):
    """softmax_cross_entropy_with_logits(logits, one_hot(x))  (:131-138)."""
    a0, _, z0 = _cat_parts(logits)
    x = np.asarray(x).reshape(logits.shape[:-1]).astype(np.int64)
    return np.log(z0[..., 0]) - np.take_along_axis(a0, x[..., None], axis=-1)[..., 0]


def cat_logp(logits, x):
    return -cat_neglogp(logits, x)                                # :25-26


def cat_entropy(logits):
    a0, ea0, z0 = _cat_parts(logits)
    return np.sum(ea0 / z0 * (np.log(z0) - a0), axis=-1)          # :148-153


def cat_kl(logits_p, logits_q):
    a0, ea0, z0 = _cat_parts(logits_p)
    a1, _, z1 = _cat_parts(logits_q)
    return np.sum(ea0 / z0 * (a0 - np.log(z0) - a1 + np.log(z1)), axis=-1)   # :139-147


def cat_mode(logits):
    return np.argmax(logits, axis=-1)                             # :129-130


def cat_sample(logits, u):
    return np.argmax(logits - np.log(-np.log(u)), axis=-1)        # :154-156 (Gumbel max)


def _ppo_losses_categorical(v, logits, v_old, logits_old, act, adv, ret, l_mul, hp):
    B = v.shape[0]
    c = hp.CLIP_PARAM * l_mul
    x = np.asarray(act).reshape(B).astype(np.int64)
    a0, ea0, z0 = _cat_parts(logits)
    p = ea0 / z0
    logp_all = a0 - np.log(z0)
    ent = cat_entropy(logits)                                     # PPO.py:29
    ratio = np.exp(cat_logp(logits, x) - cat_logp(logits_old, x))  # :31
    surr1 = ratio * adv
    surr2 = np.clip(ratio, 1.0 - c, 1.0 + c) * adv
    policy_loss = -np.mean(np.minimum(surr1, surr2))
    entropy_loss = (-hp.ENTCOEFF) * np.mean(ent)
    vf1 = np.square(v - ret)
    dv = v - v_old
    vclip = v_old + np.clip(dv, -c, c)
    vf2 = np.square(vclip - ret)
    value_loss = hp.VCOEFF * np.mean(np.maximum(vf1, vf2))
    total = policy_loss + entropy_loss + value_loss
    g_logp = np.where(surr1 <= surr2, -adv / B, 0.0) * ratio
    onehot = (x[:, None] == np.arange(logits.shape[-1])).astype(logits.dtype)
    # d logp(a)/d logit_k = [k == a] - p_k ;  d H/d logit_k = -p_k (log p_k + H)
    d_logits = g_logp[:, None] * (onehot - p) + (hp.ENTCOEFF / B) * p * (logp_all + ent[:, None])
    in_clip = (dv >= -c) & (dv <= c)
    d_v = (hp.VCOEFF / B) * np.where(vf1 >= vf2, 2.0 * (v - ret), 2.0 * (vclip - ret) * in_clip)
    d_out = np.concatenate([d_v[:, None], d_logits], axis=1)
    aux = dict(mean_kl=float(np.mean(cat_kl(logits_old, logits))), clip_frac=float(np.mean(surr1 > surr2)),
               vclip_frac=float(np.mean((vf1 < vf2) & ~in_clip)), ratio=ratio)
    return np.array([policy_loss, value_loss, entropy_loss, total]), d_out, aux


# --------------------------------------------------------------------------
# PPO loss + closed-form gradient  (PPO.py:19,29-40,46)
# --------------------------------------------------------------------------
@dataclass
class HParams:
    """main.py:15-19 defaults."""
    ENTCOEFF: float = 0.01
    VCOEFF: float = 0.5
    CLIP_PARAM: float = 0.2
    GAMMA: float = 0.99
    LAM: float = 0.95
    LEARNING_RATE: float = 2e-5


def ppo_losses_from_heads(v, flat, v_old, flat_old, act, adv, ret, l_mul, hp: HParams, pd: str = "gauss"):
    """Loss values and d(total)/d[v, mean, logstd] (or d/d[v, logits] for pd="cat") per row (canonical per-sample mode).

    TF tie rules: minimum->first arg iff x<=y; maximum->first iff
    x>=y; clip_by_value passes gradient iff lo<=x<=hi.
    """
    if pd == "cat":
        return _ppo_losses_categorical(v, flat, v_old, flat_old, act, adv, ret, l_mul, hp)
    B = v.shape[0]
    A = act.shape[1]
    c = hp.CLIP_PARAM * l_mul                                     # PPO.py:19
    mean, logstd = dg_split(flat)
    std = np.exp(logstd)
    z = (act - mean) / std
    ent = dg_entropy(flat)                                        # PPO.py:29
    logp = dg_logp(flat, act)
    logp_old = dg_logp(flat_old, act)
    ratio = np.exp(logp - logp_old)                               # PPO.py:31
    surr1 = ratio * adv                                           # :32
    surr2 = np.clip(ratio, 1.0 - c, 1.0 + c) * adv                # :33
    policy_loss = -np.mean(np.minimum(surr1, surr2))              # :34
    entropy_loss = (-hp.ENTCOEFF) * np.mean(ent)                  # :30,35
    vf1 = np.square(v - ret)                                      # :36
    dv = v - v_old
    vclip = v_old + np.clip(dv, -c, c)                            # :37
    vf2 = np.square(vclip - ret)                                  # :38
    value_loss = hp.VCOEFF * np.mean(np.maximum(vf1, vf2))        # :39
    total = policy_loss + entropy_loss + value_loss               # :40

    # ---- closed-form backward (PPO.py:46 = TF autodiff of the above) ----
    g_ratio = np.where(surr1 <= surr2, -adv / B, 0.0)
    g_logp = g_ratio * ratio
    d_mean = g_logp[:, None] * z / std
    d_logstd = g_logp[:, None] * (np.square(z) - 1.0) - hp.ENTCOEFF / B
    in_clip = (dv >= -c) & (dv <= c)
    d_v = (hp.VCOEFF / B) * np.where(vf1 >= vf2, 2.0 * (v - ret),
                                     2.0 * (vclip - ret) * in_clip)
    d_out = np.concatenate([d_v[:, None], d_mean, d_logstd], axis=1)   # [B, 1+2A]

    aux = dict(
        mean_kl=float(np.mean(dg_kl(flat_old, flat))),            # distributions.py:199-201 (diagnostic)
        clip_frac=float(np.mean(surr1 > surr2)),
        vclip_frac=float(np.mean((vf1 < vf2) & ~in_clip)),   # clipped value branch selected with the clip active
        ratio=ratio,
    )
    losses = np.array([policy_loss, value_loss, entropy_loss, total])
    return losses, d_out, aux


def mlp_backward(theta, obs, hs, d_out, lay: Layout, dtype=np.float64, q=_ident):
    """Dense backward: dW = h^T d, db = sum_b d, dh = d W^T * [h>0]; flat grad in layout order."""
    p = unflatten(np.asarray(theta, dtype), lay)
    g = np.zeros(lay.n_params, dtype)
    gp = unflatten(g, lay)
    L = len(lay.hidden)
    hL = hs[-1] if L else np.asarray(obs, dtype)
    d_v = d_out[:, :1]
    d_p = d_out[:, 1:]
    gp["Wv"][...] = q(hL).T @ q(d_v)
    gp["bv"][...] = d_v.sum(0)
    gp["Wp"][...] = q(hL).T @ q(d_p)
    gp["bp"][...] = d_p.sum(0)
    dh = q(d_v) @ q(p["Wv"]).T + q(d_p) @ q(p["Wp"]).T
    for l in range(L, 0, -1):
        dh = dh * (hs[l - 1] > 0)
        h_in = hs[l - 2] if l >= 2 else np.asarray(obs, dtype)
        gp[f"W{l}"][...] = q(h_in).T @ q(dh)
        gp[f"b{l}"][...] = dh.sum(0)
        if l > 1:
            dh = q(dh) @ q(p[f"W{l}"]).T
    return g


def ppo_loss_grad(theta_pi, theta_old, obs, act, adv, ret, l_mul, lay: Layout,
                  hp: HParams = HParams(), dtype=np.float64, q=_ident, literal_groups=None):
    """One tower: pi fwd + oldpi fwd + loss + backward  (PPO.py:21-22,29-40,46).

    Returns (losses[4] = policy, value, entropy, total ; flat_grad ; aux dict).
    literal_groups = G: the as-written pairwise value loss (quirk D7) over G time-major rollouts instead of the
    canonical per-sample one (value_literal_pairwise_grad).
    """
    obs = np.asarray(obs, dtype); act = np.asarray(act, dtype if lay.pd == "gauss" else np.int64)
    adv = np.asarray(adv, dtype).reshape(-1); ret = np.asarray(ret, dtype).reshape(-1)
    v, flat, hs = mlp_forward(theta_pi, obs, lay, dtype, q)
    v_old, flat_old, _ = mlp_forward(theta_old, obs, lay, dtype, q)
    losses, d_out, aux = ppo_losses_from_heads(v, flat, v_old, flat_old, act, adv, ret,
                                               dtype(l_mul), hp, lay.pd)
    if literal_groups:
        vl, d_v, clipped = value_literal_pairwise_grad(v, v_old, ret, float(l_mul), int(literal_groups), hp)
        losses = np.array([losses[0], vl, losses[2], losses[0] + losses[2] + vl])
        d_out = d_out.copy(); d_out[:, 0] = d_v
        aux["vclip_frac"] = float(clipped)
    grad = mlp_backward(theta_pi, obs, hs, d_out.astype(dtype), lay, dtype, q)
    aux.update(d_out=d_out, v=v, flat=flat, v_old=v_old, flat_old=flat_old, hs=hs)
    return losses, grad, aux


def value_literal_pairwise_grad(v, v_old, ret, l_mul, groups=1, hp: HParams = HParams()):
    """Quirk D7 as a trainable mode: PPO.py:36-39 as written pairs every value v_i with every return R_j of its tower
    ([B,1,1] - [B,1] -> [B,B,1]).  A tower of the reference is ONE rollout; a time-major batch of `groups` rollouts
    (row b = t*groups + n) is `groups` towers, whose losses / gradients are averaged (PPO.py:55-65).
    Returns (value_loss, d value_loss / d v [B], fraction of pairs on the clipped branch with the clip active);
    TF tie rules: maximum -> first arg iff x >= y, clip_by_value passes gradient iff lo <= x <= hi."""
    c = hp.CLIP_PARAM * l_mul
    v = np.asarray(v, np.float64).reshape(-1); v_old = np.asarray(v_old, np.float64).reshape(-1)
    R = np.asarray(ret, np.float64).reshape(-1)
    B = v.shape[0]; Tn = B // groups
    loss = 0.0; clipped = 0.0
    d_v = np.zeros(B)
    for g in range(groups):
        idx = np.arange(Tn) * groups + g
        vi = v[idx][:, None]; vo = v_old[idx][:, None]; Rj = R[idx][None, :]
        dv = vi - vo
        vclip = vo + np.clip(dv, -c, c)
        vf1 = np.square(vi - Rj); vf2 = np.square(vclip - Rj)
        in_clip = (dv >= -c) & (dv <= c)
        loss += hp.VCOEFF * np.mean(np.maximum(vf1, vf2)) / groups
        clipped += np.mean((vf1 < vf2) & ~in_clip) / groups
        d_v[idx] = (hp.VCOEFF / (groups * Tn * Tn)) * np.sum(
            np.where(vf1 >= vf2, 2.0 * (vi - Rj), 2.0 * (vclip - Rj) * in_clip), axis=1)
    return loss, d_v, clipped


def value_loss_literal_pairwise(v, v_old, ret, l_mul, hp: HParams = HParams()):
    """Quirk D7: PPO.py:36-39 as written broadcasts [B,1,1]-[B,1] -> [B,B,1] (all pairs)."""
    c = hp.CLIP_PARAM * l_mul
    v = np.asarray(v, np.float64).reshape(-1, 1)
    v_old = np.asarray(v_old, np.float64).reshape(-1, 1)
    R = np.asarray(ret, np.float64).reshape(1, -1)
    vf1 = np.square(v - R)
    vf2 = np.square(v_old + np.clip(v - v_old, -c, c) - R)
    return hp.VCOEFF * np.mean(np.maximum(vf1, vf2))


# --------------------------------------------------------------------------
# GAE  (Worker.py:82-92)
# --------------------------------------------------------------------------
GAE_REFERENCE = 0   # mask = 1-done[t+1], bootstrap = predv[T-1]*(1-done[T-1])   (quirk D8)
GAE_STANDARD = 1    # mask = 1-done[t],   bootstrap = last_v*(1-last_done) given explicitly


def gae_rollout(r, predv, done, gamma, lam, mode=GAE_REFERENCE, last_v=None, last_done=None,
                normalize=True):
    """One rollout (python loop, fp64 scalars, fp32 store -- as Worker.py:82-92)."""
    T = len(r)
    r = np.asarray(r, np.float64); predv32 = np.asarray(predv, np.float32)
    pv = predv32.astype(np.float64); done = np.asarray(done, np.float64)
    if mode == GAE_REFERENCE:
        d_ext = np.append(done, 0.0)                              # :82
        v_ext = np.append(pv, pv[-1] * (1.0 - done[-1]))          # :83
        nt = 1.0 - d_ext[1:]
    else:
        lv = 0.0 if last_v is None else float(last_v)
        ld = 0.0 if last_done is None else float(last_done)
        v_ext = np.append(pv, lv * (1.0 - ld))
        nt = 1.0 - done
    adv = np.empty(T, np.float32)                                 # :85
    last = 0.0
    for t in reversed(range(T)):                                  # :86-90
        delta = r[t] + gamma * v_ext[t + 1] * nt[t] - v_ext[t]
        adv[t] = last = delta + gamma * lam * nt[t] * last
    etr = adv.astype(np.float64) + pv                             # :91 (fp64; fed to an fp32 placeholder)
    if normalize:
        adv = (adv - adv.mean()) / adv.std()                      # :92 (fp32, ddof 0, no eps)
    return adv, etr.astype(np.float32)


def gae_batch(rew, predv, done, gamma, lam, mode=GAE_REFERENCE, last_v=None, last_done=None,
              normalize=True):
    """[T,N] time-major; every column is one rollout (one reference Worker)."""
    T, N = rew.shape
    adv = np.empty((T, N), np.float32); etr = np.empty((T, N), np.float32)
    for n in range(N):
        a, e = gae_rollout(rew[:, n], predv[:, n], done[:, n], gamma, lam, mode,
                           None if last_v is None else last_v[n],
                           None if last_done is None else last_done[n], normalize)
        adv[:, n] = a; etr[:, n] = e
    return adv, etr


def gae_batch_vectorized(rew, predv, done, gamma, lam, mode=GAE_REFERENCE, last_v=None,
                         last_done=None, normalize=True):
    """Same numbers, numpy-vectorised over N (used for the CPU-baseline timing and big parity)."""
    T, N = rew.shape
    r = rew.astype(np.float64); pv = predv.astype(np.float32).astype(np.float64)
    d = done.astype(np.float64)
    if mode == GAE_REFERENCE:
        v_next = np.vstack([pv[1:], (pv[-1] * (1.0 - d[-1]))[None]])
        nt = np.vstack([1.0 - d[1:], np.ones((1, N))])
    else:
        lv = np.zeros(N) if last_v is None else np.asarray(last_v, np.float64)
        ld = np.zeros(N) if last_done is None else np.asarray(last_done, np.float64)
        v_next = np.vstack([pv[1:], (lv * (1.0 - ld))[None]])
        nt = 1.0 - d
    adv = np.empty((T, N), np.float32)
    last = np.zeros(N)
    for t in range(T - 1, -1, -1):
        delta = r[t] + gamma * v_next[t] * nt[t] - pv[t]
        last = delta + gamma * lam * nt[t] * last
        adv[t] = last
    etr = (adv.astype(np.float64) + pv).astype(np.float32)
    if normalize:
        adv = (adv - adv.mean(axis=0, dtype=np.float32)) / adv.std(axis=0, dtype=np.float32)
    return adv.astype(np.float32), etr


# --------------------------------------------------------------------------
# tower mean + Adam + schedule  (PPO.py:18-20,53,55-65; Worker.py:77-80)
# --------------------------------------------------------------------------
def tower_mean(grads: List[np.ndarray]) -> np.ndarray:
    """PPO.py:55-65: expand_dims/concat/reduce_mean over towers."""
    return np.mean(np.stack([np.asarray(g) for g in grads], axis=0), axis=0)


def adam_tf(theta, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8, dtype=np.float64):
    """tf.train.AdamOptimizer (PPO.py:20,53), TF-1.x epsilon-hat form; t is the step AFTER increment."""
    theta = np.asarray(theta, dtype); g = np.asarray(g, dtype)
    # TF casts beta1/beta2 to the variable dtype (fp32) before forming (1 - beta): keep that rounding
    b1 = float(np.float32(beta1)); b2 = float(np.float32(beta2))
    omb1 = float(np.float32(1) - np.float32(beta1)); omb2 = float(np.float32(1) - np.float32(beta2))
    m = b1 * np.asarray(m, dtype) + omb1 * g
    v = b2 * np.asarray(v, dtype) + omb2 * g * g
    lr_t = dtype(lr) * np.sqrt(1 - dtype(beta2) ** t) / (1 - dtype(beta1) ** t)
    theta = theta - lr_t * m / (np.sqrt(v) + eps)
    return theta, m, v


# --------------------------------------------------------------------------
# epoch / minibatch scheduler.  Not in the reference (Chief.py:58-64 feeds whole rollouts
# UPDATE_STEPS times); restates this repo's csrc/perm.cuh so that the device scheduler has a bit-exact checker.
# --------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def _perm_mix(x, key):
    x = x ^ np.uint32(key)
    x = x * np.uint32(0x9E3779B1); x = x ^ (x >> np.uint32(15))
    x = x * np.uint32(0x85EBCA77); x = x ^ (x >> np.uint32(13))
    x = x * np.uint32(0xC2B2AE3D); x = x ^ (x >> np.uint32(16))
    return x


def minibatch_perm(n: int, seed: int, epoch: int) -> np.ndarray:
    """Keyed bijection of [0, n): 4-round (unbalanced) Feistel over max(2, ceil(log2 n)) bits + cycle walking."""
    if n == 0:
        return np.zeros(0, np.int64)
    bits = 2
    while bits < 31 and (1 << bits) < n:
        bits += 1
    s = _splitmix64((seed ^ ((epoch * 0xD1B54A32D192ED03) & _M64)) & _M64)
    keys = []
    for _ in range(4):
        s = _splitmix64(s)
        keys.append(s >> 32)
    a0 = bits // 2; b0 = bits - a0
    out = np.arange(n, dtype=np.uint32)
    todo = np.ones(n, bool)
    with np.errstate(over="ignore"):
        while todo.any():
            i = out[todo]
            la, lb = a0, b0
            L = i >> np.uint32(b0); R = i & np.uint32((1 << b0) - 1)
            for r in range(4):
                nl = R; nr = L ^ (_perm_mix(R, keys[r]) & np.uint32((1 << la) - 1))
                L, R = nl, nr
                la, lb = lb, la
            i = (L << np.uint32(lb)) | R
            out[todo] = i
            todo[todo] = i >= n
    return out.astype(np.int64)


def update_epochs(theta_pi, theta_old, buf, gamma, lam, l_mul, lay: Layout, lr, n_epochs=1, minibatch_rows=0,
                  shuffle=False, seed=0, step0=1, hp: HParams = HParams(), dtype=np.float64, q=_ident):
    """GAE once (Worker.py:82-92), then epochs x minibatches of [loss_grad (PPO.py:21-46) -> Adam (PPO.py:20,53)];
    n_epochs=UPDATE_STEPS with one whole-buffer minibatch is Chief.py:64.  Returns (theta, m, v, losses[k,4])."""
    adv, ret = gae_batch_vectorized(buf["rew"], buf["predv"], buf["done"], gamma, lam)
    adv = adv.reshape(-1); ret = ret.reshape(-1)
    B = adv.size
    mb = B if not minibatch_rows else min(minibatch_rows, B)
    theta = np.asarray(theta_pi, np.float64).copy()
    m = np.zeros_like(theta); v = np.zeros_like(theta)
    step = step0; out = []
    for e in range(n_epochs):
        order = minibatch_perm(B, seed, e) if shuffle else np.arange(B)
        for r0 in range(0, B, mb):
            idx = order[r0:r0 + mb]
            losses, g, _ = ppo_loss_grad(theta.astype(np.float32), theta_old, buf["obs"][idx], buf["act"][idx], adv[idx],
                                         ret[idx], l_mul, lay, hp, dtype, q)
            theta, m, v = adam_tf(theta.astype(np.float32), g.astype(np.float32), m, v, step, lr)
            out.append(losses); step += 1
    return theta, m, v, np.array(out)


def lr_multiplier(schedule: str, cur_ep: int, epoch_max: int) -> float:
    """Worker.py:77-80."""
    if schedule == "constant":
        return 1.0
    if schedule == "linear":
        return max(1.0 - float(cur_ep) / epoch_max, 0)
    raise NotImplementedError(schedule)


# --------------------------------------------------------------------------
# synthetic frozen buffers; shared by tests, smoke and bench
# --------------------------------------------------------------------------
def make_buffer(T, N, obs_dim, act_dim, seed=1234, pd="gauss"):
    """pd="cat": act is the reference's int32 [B,1] placeholder (PPO.py:26), classes uniform in [0, act_dim)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    B = T * N
    obs = torch.randn(B, obs_dim, generator=g, dtype=torch.float32).numpy()
    if pd == "cat":
        act = torch.randint(0, act_dim, (B, 1), generator=g, dtype=torch.int32).numpy()
    else:
        act = torch.randn(B, act_dim, generator=g, dtype=torch.float32).numpy()
    rew = torch.randn(T, N, generator=g, dtype=torch.float32).numpy()
    done = (torch.rand(T, N, generator=g) < (1.0 / 200.0)).to(torch.float32).numpy()
    predv = torch.randn(T, N, generator=g, dtype=torch.float32).numpy()
    return dict(obs=obs, act=act, rew=rew, done=done, predv=predv)


def make_params_pair(lay: Layout, trunk_std=1.0, head_std=0.1, delta=0.02, seed_old=7, seed_pi=8):
    """theta_old ~ normc(trunk_std / head_std), theta_pi = theta_old + delta*N(0,1)  (pi != oldpi)."""
    theta_old = init_params(lay, np.random.default_rng(seed_old), trunk_std, head_std)
    noise = np.random.default_rng(seed_pi).standard_normal(lay.n_params).astype(np.float32)
    theta_pi = (theta_old + delta * noise).astype(np.float32)
    return theta_pi, theta_old


# --------------------------------------------------------------------------
# the reference-equivalent full update on CPU (bench --impl reference / cpu_baseline)
# --------------------------------------------------------------------------
def cpu_update_torch(theta_pi, theta_old, buf, T, N, lay: Layout, hp: HParams, l_mul, n_towers,
                     threads=None):
    """fp32 PyTorch-CPU restatement structured like the reference graph: ``n_towers`` worker
    towers (PPO.py:21-46 each: pi fwd, oldpi fwd, loss, autodiff backward), per-variable tower
    mean (PPO.py:55-65), GAE per rollout (Worker.py:82-92, vectorised over columns).
    Returns (losses_mean[4], mean_grad)."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    adv, etr = gae_batch_vectorized(buf["rew"], buf["predv"], buf["done"], hp.GAMMA, hp.LAM)
    obs = torch.from_numpy(buf["obs"]); act = torch.from_numpy(buf["act"])
    advt = torch.from_numpy(adv.reshape(-1)); rett = torch.from_numpy(etr.reshape(-1))
    B = obs.shape[0]

    def net(theta, x):
        h = x
        for name, shape, off in lay.entries:
            n = int(np.prod(shape))
            w = theta[off:off + n].view(*shape)
            if name.startswith("W") and name not in ("Wv", "Wp"):
                W = w
            elif name.startswith("b") and name not in ("bv", "bp"):
                h = torch.relu(h @ W + w)
            elif name == "Wv":
                Wv = w
            elif name == "bv":
                v = (h @ Wv + w)[:, 0]
            elif name == "Wp":
                Wp = w
            else:
                flat = h @ Wp + w
        return v, flat

    def logp(flat, x):
        A = x.shape[1]
        mean, logstd = flat[:, :A], flat[:, A:]
        return -(0.5 * (((x - mean) / torch.exp(logstd)) ** 2).sum(-1) + 0.5 * LOG_2PI * A
                 + logstd.sum(-1))

    t_old = torch.from_numpy(np.asarray(theta_old, np.float32))
    grads, losses = [], []
    c = hp.CLIP_PARAM * l_mul
    rows = np.array_split(np.arange(B), n_towers)
    for idx in rows:
        sl = slice(int(idx[0]), int(idx[-1]) + 1)
        t_pi = torch.from_numpy(np.asarray(theta_pi, np.float32)).clone().requires_grad_(True)
        v, flat = net(t_pi, obs[sl])
        with torch.no_grad():
            v_old, flat_old = net(t_old, obs[sl])
        A = act.shape[1]
        ent = (flat[:, A:] + 0.5 * math.log(2 * math.pi * math.e)).sum(-1)
        ratio = torch.exp(logp(flat, act[sl]) - logp(flat_old, act[sl]))
        s1 = ratio * advt[sl]
        s2 = torch.clamp(ratio, 1.0 - c, 1.0 + c) * advt[sl]
        pl = -torch.minimum(s1, s2).mean()
        el = -hp.ENTCOEFF * ent.mean()
        vf1 = (v - rett[sl]) ** 2
        vf2 = (v_old + torch.clamp(v - v_old, -c, c) - rett[sl]) ** 2
        vl = hp.VCOEFF * torch.maximum(vf1, vf2).mean()
        tot = pl + el + vl
        tot.backward()
        grads.append(t_pi.grad.numpy())
        losses.append([pl.item(), vl.item(), el.item(), tot.item()])
    return np.mean(np.array(losses), axis=0), tower_mean(grads)


# --------------------------------------------------------------------------
# branch-robust synthetic buffers for parity tests
# --------------------------------------------------------------------------
def sanitize_obs(theta_pi, obs, lay: Layout, margin=2e-5, seed=98):
    """ReLU is the other discontinuity of the backward pass: tf.nn.relu's gradient is [h > 0], so a
    pre-activation within fp32 rounding of zero flips one element of dh between the fp32 device
    result and the fp64 oracle.  Resample the SYNTHETIC observation rows that have any pi
    pre-activation with |z| < margin until none is left.  Returns (obs, n_resampled)."""
    rng = np.random.default_rng(seed)
    obs = np.array(obs, np.float32, copy=True)
    p = unflatten(np.asarray(theta_pi, np.float64), lay)
    rows = np.arange(obs.shape[0]); total = 0
    for _ in range(200):
        h = obs[rows].astype(np.float64)
        bad = np.zeros(len(rows), bool)
        for l in range(len(lay.hidden)):
            z = h @ p[f"W{l + 1}"] + p[f"b{l + 1}"]
            bad |= (np.abs(z) < margin).any(axis=1)
            h = np.maximum(z, 0)
        rows = rows[bad]
        if len(rows) == 0:
            return obs, total
        total += len(rows)
        obs[rows] = rng.standard_normal((len(rows), obs.shape[1])).astype(np.float32)
    raise RuntimeError("sanitize_obs did not converge")


def sanitize_rows(theta_pi, theta_old, obs, act, adv, ret, l_mul, lay: Layout, hp: HParams = HParams(),
                  margin=1e-4, seed=99):
    """The PPO loss is discontinuous in its inputs where tf.minimum / tf.maximum / clip switch
    branch (PPO.py:33-39).  A row that sits within fp32 rounding of such a switch makes the
    fp32 device result and the fp64 oracle legitimately pick different branches (a 1/sqrt(B)
    relative change of the gradient), which says nothing about parity.  This helper edits the
    SYNTHETIC act / ret of the few rows within ``margin`` of a switch so that none is left:
      * |ratio - (1 +- c)| < margin                -> resample the action row;
      * ||v - v_old| - c| < margin (clip edge)     -> put the return on the far side of v_old, where
                                                      both branches have the same gradient 2(v-R);
      * outside the clip and |vf1 - vf2| < margin  -> resample the return.
    Returns (act, ret, n_edited)."""
    rng = np.random.default_rng(seed)
    cat = lay.pd == "cat"
    act = np.array(act, np.int32 if cat else np.float32, copy=True); ret = np.array(ret, np.float32, copy=True).reshape(-1)
    v, flat, _ = mlp_forward(theta_pi, obs, lay)
    v_old, flat_old, _ = mlp_forward(theta_old, obs, lay)
    c = hp.CLIP_PARAM * l_mul
    dv = v - v_old
    edited = 0
    for _ in range(50):
        a64 = act.astype(np.float64); r64 = ret.astype(np.float64)
        ratio = (np.exp(cat_logp(flat, act) - cat_logp(flat_old, act)) if cat else
                 np.exp(dg_logp(flat, a64) - dg_logp(flat_old, a64)))
        bad_r = (np.abs(ratio - (1 - c)) < margin) | (np.abs(ratio - (1 + c)) < margin)
        edge = np.abs(np.abs(dv) - c) < margin
        vclip = v_old + np.clip(dv, -c, c)
        vf1 = np.square(v - r64); vf2 = np.square(vclip - r64)
        on_far_side = np.sign(dv) * (v_old - r64) > 0.5
        bad_edge = edge & ~on_far_side
        bad_v = (np.abs(dv) > c) & ~edge & (np.abs(vf1 - vf2) < margin * (1 + vf1))
        n = int(bad_r.sum() + bad_edge.sum() + bad_v.sum())
        if n == 0:
            break
        edited += n
        act[bad_r] = (rng.integers(0, lay.act_dim, (int(bad_r.sum()), 1)).astype(np.int32) if cat else
                      rng.standard_normal((int(bad_r.sum()), act.shape[1])).astype(np.float32))
        sgn = np.where(dv[bad_edge] >= 0, 1.0, -1.0)
        ret[bad_edge] = (v_old[bad_edge] - sgn * (1.0 + np.abs(rng.standard_normal(int(bad_edge.sum()))))).astype(np.float32)
        ret[bad_v] = rng.standard_normal(int(bad_v.sum())).astype(np.float32)
    else:
        raise RuntimeError("sanitize_rows did not converge")
    return act, ret, edited


def sanitize_for_update(theta_pi, theta_old, obs, act, adv, ret, l_mul, lay: Layout, hp: HParams = HParams(),
                        margin=1e-4, relu_margin=2e-5, seed=97):
    """Branch-robust inputs for the WHOLE-UPDATE entry points (dppo_update / dppo_update_host), which compute
    adv / ret themselves from rew / predv / done: those two are therefore fixed here and only the synthetic obs / act
    rows are edited.  A row is resampled until it is (i) not within relu_margin of a ReLU switch of pi
    (sanitize_obs), (ii) not within margin of a ratio-clip switch (new action row), (iii) not within margin of a
    value-clip / max switch (new observation row: predv is an independent input of the synthetic buffer, so v and
    v_old move and the GAE outputs do not).  Same criteria as sanitize_rows.  Returns (obs, act, n_edited)."""
    rng = np.random.default_rng(seed)
    obs = np.array(obs, np.float32, copy=True)
    act = np.array(act, np.float32, copy=True)
    ret = np.asarray(ret, np.float64).reshape(-1)
    c = hp.CLIP_PARAM * l_mul
    p = unflatten(np.asarray(theta_pi, np.float64), lay)
    rows = np.arange(obs.shape[0]); edited = 0
    for _ in range(200):
        o = obs[rows].astype(np.float64)
        h = o; bad_relu = np.zeros(len(rows), bool)
        for l in range(len(lay.hidden)):
            z = h @ p[f"W{l + 1}"] + p[f"b{l + 1}"]
            bad_relu |= (np.abs(z) < relu_margin).any(axis=1)
            h = np.maximum(z, 0)
        v, flat, _ = mlp_forward(theta_pi, obs[rows], lay)
        v_old, flat_old, _ = mlp_forward(theta_old, obs[rows], lay)
        a64 = act[rows].astype(np.float64); r64 = ret[rows]
        ratio = np.exp(dg_logp(flat, a64) - dg_logp(flat_old, a64))
        bad_r = (np.abs(ratio - (1 - c)) < margin) | (np.abs(ratio - (1 + c)) < margin)
        dv = v - v_old
        edge = np.abs(np.abs(dv) - c) < margin
        vclip = v_old + np.clip(dv, -c, c)
        vf1 = np.square(v - r64); vf2 = np.square(vclip - r64)
        on_far_side = np.sign(dv) * (v_old - r64) > 0.5
        bad_v = (edge & ~on_far_side) | ((np.abs(dv) > c) & ~edge & (np.abs(vf1 - vf2) < margin * (1 + vf1)))
        bad_obs = bad_relu | bad_v
        bad = bad_obs | bad_r
        if not bad.any():
            return obs, act, edited
        edited += int(bad.sum())
        obs[rows[bad_obs]] = rng.standard_normal((int(bad_obs.sum()), obs.shape[1])).astype(np.float32)
        act[rows[bad_r]] = rng.standard_normal((int(bad_r.sum()), act.shape[1])).astype(np.float32)
        rows = rows[bad]
    raise RuntimeError("sanitize_for_update did not converge")


def make_case(T, N, obs_dim, hidden, act_dim, seed=1234, l_mul=0.7, trunk_std=1.0, head_std=0.1, delta=0.02,
              hp: HParams = HParams(), pd="gauss"):
    """Frozen synthetic update case: buffer, GAE'd advantages/returns (oracle),
    pi != oldpi parameters, with branch-ambiguous rows removed.  Everything fp32."""
    lay = make_layout(obs_dim, hidden, act_dim, pd)
    buf = make_buffer(T, N, obs_dim, act_dim, seed, pd)
    th_pi, th_old = make_params_pair(lay, trunk_std, head_std, delta)
    adv, ret = gae_batch_vectorized(buf["rew"], buf["predv"], buf["done"], hp.GAMMA, hp.LAM)
    buf["obs"], n_obs = sanitize_obs(th_pi, buf["obs"], lay)
    act, ret2, edited = sanitize_rows(th_pi, th_old, buf["obs"], buf["act"], adv.reshape(-1), ret.reshape(-1),
                                      l_mul, lay, hp)
    return dict(lay=lay, buf=buf, theta_pi=th_pi, theta_old=th_old, adv=adv.reshape(-1).astype(np.float32),
                ret=ret2.astype(np.float32), act=act, l_mul=l_mul, edited=edited + n_obs, T=T, N=N)
