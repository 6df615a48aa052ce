#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE in this container.

TEST INFRASTRUCTURE.  Run here (``/root/reference`` exists only in the authoring container):

    python oracle/make_golden.py

Nothing from /root/reference is copied into the repo: the script reads the reference files at
run time, executes them (whole modules, or the cited line ranges for code that lives inside
functions that cannot run as committed -- SURVEY Appendix A), and stores only numeric
inputs/outputs.

  gae_ref.npz     Worker.py:82-92 executed verbatim (pure numpy -> genuinely the reference).
  distgauss_ref.npz  Others/distributions.py DiagGaussianPd methods, executed over
                     oracle/tf_numpy_shim.py (TensorFlow is not installed).
  model_fc_ref.npz   Model.py  Model().FC(...)  (+ tf_util.normc_initializer) over the shim.
  ppo_loss_ref.npz   PPO.py:29-40 executed over the shim, on canonical rank ([B,1]) inputs and
                     on the literal rank-3 value head (quirk D7, pairwise broadcast).
"""
import importlib.util
import os
import sys
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, HERE)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    import tf_numpy_shim
    tf = tf_numpy_shim.install()
    # the reference imports itself as package ``RL``
    RL = types.ModuleType("RL"); RL.__path__ = []
    Others = types.ModuleType("RL.Others"); Others.__path__ = []
    sys.modules["RL"] = RL; sys.modules["RL.Others"] = Others
    # gym is not installed: the two space classes make_pdtype() dispatches on
    gym = types.ModuleType("gym"); spaces = types.ModuleType("gym.spaces")

    class Box:
        def __init__(self, shape): self.shape = tuple(shape)

    class Discrete:
        def __init__(self, n): self.n = n

    class _Other:
        pass
    spaces.Box, spaces.Discrete = Box, Discrete
    spaces.MultiDiscrete = spaces.MultiBinary = _Other
    gym.spaces = spaces
    sys.modules["gym"] = gym; sys.modules["gym.spaces"] = spaces
    U = _load("RL.Others.tf_util", os.path.join(REF, "Others", "tf_util.py"))
    Others.tf_util = U
    D = _load("RL.Others.distributions", os.path.join(REF, "Others", "distributions.py"))
    Others.distributions = D
    M = _load("RL.Model", os.path.join(REF, "Model.py"))
    return tf, U, D, M, spaces


def ref_lines(fname, lo, hi):
    with open(os.path.join(REF, fname)) as f:
        lines = f.readlines()[lo - 1:hi]
    return textwrap.dedent("".join(lines))


# ---------------------------------------------------------------- GAE
def gae_cases():
    rng = np.random.default_rng(20260921)
    cases = []
    # SURVEY Appendix C known-answer case
    cases.append((np.array([1, .5, -1, 2, 0, 1.5]), np.array([.1, -.2, .3, .4, -.5, .6]),
                  np.array([0, 0, 1, 0, 0, 0], bool)))
    for T, p in [(200, 1 / 50), (200, 1 / 50), (200, 0.0), (200, 1 / 10), (2, 0.0), (7, 0.5),
                 (2048, 1 / 200), (2048, 1 / 200), (33, 1.0), (513, 1 / 20)]:
        r = rng.standard_normal(T).astype(np.float32).astype(np.float64)   # fp32-representable rewards
        v = rng.standard_normal(T)
        d = rng.random(T) < p
        cases.append((r, v, d))
    # done on the very last step (bootstrap masked) and on the first step
    r = rng.standard_normal(50).astype(np.float32).astype(np.float64); v = rng.standard_normal(50)
    d = np.zeros(50, bool); d[-1] = True
    cases.append((r, v, d))
    d = np.zeros(50, bool); d[0] = True
    cases.append((r.copy(), v.copy(), d))
    return cases


def run_gae_reference(src, r, v, d, gamma, lam, numpy1_promotion=True):
    """Worker.py:82-92 with the buffers as the rollout loop leaves them (Worker.py:52-56,70-75).

    The reference predates NumPy 2: with NumPy 1.x value-based promotion ``python_float *
    np.float32_scalar`` is float64, so Worker.py:88-90 ran in fp64.  NumPy 2.3 (this container,
    NEP 50) keeps such products in fp32.  ``numpy1_promotion=True`` feeds the fp32-representable
    value predictions as python floats, which makes NumPy 2 reproduce the fp64 arithmetic of
    the reference's era bit for bit; ``False`` feeds np.float32 scalars as sess.run returned
    them and records what NumPy 2 literally computes (differs by <= 2 fp32 ulp)."""
    f32 = (lambda x: float(np.float32(x))) if numpy1_promotion else np.float32
    ns = dict(
        np=np,
        self=types.SimpleNamespace(GAMMA=gamma, LAM=lam),
        buffer_r=[float(x) for x in r],                       # gym rewards: python floats
        buffer_predv=[f32(x) for x in v],                     # sess.run outputs (fp32 values)
        buffer_done=[bool(x) for x in d],
    )
    ns["pred_v"] = ns["buffer_predv"][-1]                     # last loop iteration's values
    ns["done"] = ns["buffer_done"][-1]
    exec(src, ns)
    return np.asarray(ns["buffer_adv"]), np.asarray(ns["buffer_etr"])


def make_gae():
    src = ref_lines("Worker.py", 82, 92)
    out = {}
    cases = gae_cases()
    for i, (r, v, d) in enumerate(cases):
        with np.errstate(all="ignore"):
            adv, etr = run_gae_reference(src, r, v, d, 0.99, 0.95, True)
            adv2, etr2 = run_gae_reference(src, r, v, d, 0.99, 0.95, False)
        out[f"c{i}_r"] = np.asarray(r, np.float64)
        out[f"c{i}_predv"] = np.asarray(v, np.float32)
        out[f"c{i}_done"] = np.asarray(d, np.float32)
        out[f"c{i}_adv"] = adv
        out[f"c{i}_etr"] = etr
        out[f"c{i}_adv_np2"] = adv2
        out[f"c{i}_etr_np2"] = etr2
    out["n_cases"] = np.array(len(cases))
    out["gamma"] = np.array(0.99); out["lam"] = np.array(0.95)
    np.savez_compressed(os.path.join(OUT, "gae_ref.npz"), **out)
    print("gae_ref.npz:", len(cases), "cases; KAT adv =", out["c0_adv"])


# ---------------------------------------------------------------- DiagGaussianPd
def make_distgauss(D):
    out = {}
    # the reference self-check fixture (distributions.py:256) + the SURVEY Appendix-C probe points
    p = np.array([-.2, .3, .4, -.5, .1, -.5, .1, 0.8])
    x = np.array([.1, -.2, .3, .4])
    q = p + np.array([.1, -.1, .05, 0, .02, -.03, 0, .1])
    pd, pdq = D.DiagGaussianPd(p[None]), D.DiagGaussianPd(q[None])
    out["fix_p"], out["fix_x"], out["fix_q"] = p, x, q
    out["fix_entropy"] = pd.entropy(); out["fix_neglogp"] = pd.neglogp(x[None])
    out["fix_kl"] = pd.kl(pdq); out["fix_mode"] = pd.mode()
    rng = np.random.default_rng(11)
    for A in (1, 6, 17):
        fp = rng.standard_normal((64, 2 * A)) * 0.7
        fq = fp + 0.1 * rng.standard_normal((64, 2 * A))
        xs = rng.standard_normal((64, A))
        pd, pdq = D.DiagGaussianPd(fp), D.DiagGaussianPd(fq)
        pt = D.DiagGaussianPdType(A)
        assert pt.param_shape() == [2 * A] and pt.sample_shape() == [A]
        out[f"A{A}_p"], out[f"A{A}_q"], out[f"A{A}_x"] = fp, fq, xs
        out[f"A{A}_neglogp"] = pd.neglogp(xs); out[f"A{A}_logp"] = pd.logp(xs)
        out[f"A{A}_entropy"] = pd.entropy(); out[f"A{A}_kl"] = pd.kl(pdq)
        out[f"A{A}_mode"] = pd.mode()
    np.savez_compressed(os.path.join(OUT, "distgauss_ref.npz"), **out)
    print("distgauss_ref.npz: entropy", out["fix_entropy"], "neglogp", out["fix_neglogp"],
          "kl", out["fix_kl"])


# ---------------------------------------------------------------- CategoricalPd (distributions.py:124-159)
def make_categorical(tf, U, D):
    out = {}
    # the reference self-check fixture (distributions.py:260)
    p = np.array([-.2, .3, .5]); q = p + np.array([.1, -.05, .02])
    pd, pdq = D.CategoricalPd(tf.as_tensor(p[None])), D.CategoricalPd(tf.as_tensor(q[None]))
    out["fix_p"], out["fix_q"] = p, q
    out["fix_entropy"] = pd.entropy(); out["fix_kl"] = pd.kl(pdq); out["fix_mode"] = pd.mode()
    out["fix_neglogp"] = np.stack([pd.neglogp(np.array([k])) for k in range(3)])
    rng = np.random.default_rng(13)
    for n in (2, 6, 18):
        fp = rng.standard_normal((64, n)) * 1.5
        fq = fp + 0.2 * rng.standard_normal((64, n))
        xs = rng.integers(0, n, size=64)
        pd, pdq = D.CategoricalPd(tf.as_tensor(fp)), D.CategoricalPd(tf.as_tensor(fq))
        pt = D.CategoricalPdType(n)
        assert pt.param_shape() == [n] and pt.sample_shape() == []
        out[f"n{n}_p"], out[f"n{n}_q"], out[f"n{n}_x"] = fp, fq, xs
        out[f"n{n}_neglogp"] = pd.neglogp(xs); out[f"n{n}_logp"] = pd.logp(xs)
        out[f"n{n}_entropy"] = pd.entropy(); out[f"n{n}_kl"] = pd.kl(pdq)
        out[f"n{n}_mode"] = pd.mode()
    # PPO.py:29-40 over CategoricalPd heads: a is the reference's own int32 [B,1] placeholder (PPO.py:26)
    src = ref_lines("PPO.py", 29, 40)
    for tag, n, B, l_mul in (("a", 2, 64, 0.7), ("b", 6, 96, 1.0), ("c", 18, 50, 0.35)):
        flat = rng.standard_normal((B, 1, n))
        flat_old = flat + 0.2 * rng.standard_normal((B, 1, n))
        a = rng.integers(0, n, size=(B, 1))
        adv = rng.standard_normal((B, 1)); etr = rng.standard_normal((B, 1))
        v = rng.standard_normal((B, 1)); v_old = v + 0.3 * rng.standard_normal((B, 1))
        self = types.SimpleNamespace(a=a, adv=adv, etr=etr, pipredv=v)
        ns = dict(tf=tf, U=U, np=np, self=self, piPD=D.CategoricalPd(tf.as_tensor(flat)),
                  oldpiPD=D.CategoricalPd(tf.as_tensor(flat_old)), oldpredv=v_old,
                  CLIP_PARAM=0.2 * l_mul, ENTCOEFF=0.01, VCOEFF=0.5)
        exec(src, ns)
        out[f"loss_{tag}_flat"], out[f"loss_{tag}_flat_old"] = flat[:, 0], flat_old[:, 0]
        out[f"loss_{tag}_act"], out[f"loss_{tag}_adv"], out[f"loss_{tag}_etr"] = a[:, 0], adv[:, 0], etr[:, 0]
        out[f"loss_{tag}_v"], out[f"loss_{tag}_v_old"] = v[:, 0], v_old[:, 0]
        out[f"loss_{tag}_l_mul"] = np.array(l_mul)
        out[f"loss_{tag}_ratio"] = np.asarray(ns["ratio"])[:, 0]
        out[f"loss_{tag}_losses"] = np.array([self.policyLoss, self.valueLoss, self.entropyLoss, self.total_loss],
                                             np.float64)
    np.savez_compressed(os.path.join(OUT, "categorical_ref.npz"), **{k: np.asarray(v_) for k, v_ in out.items()})
    print("categorical_ref.npz: entropy", out["fix_entropy"], "neglogp", out["fix_neglogp"].ravel(), "kl", out["fix_kl"],
          "losses(a)", out["loss_a_losses"])


# ---------------------------------------------------------------- Model.FC
def make_model_fc(tf, M, spaces):
    out = {}
    for tag, obs_dim, A, B in (("pend", 3, 1, 40), ("cheetah", 17, 6, 24)):
        tf.reset_default_graph()
        np.random.seed(0)                        # normc_initializer draws from the global RNG
        s = np.random.default_rng(3).standard_normal((B, obs_dim))
        predv, pd, para = M.Model().FC("W0pi", s, spaces.Box((A,)))
        out[f"{tag}_s"] = s
        out[f"{tag}_predv"] = predv               # [B,1,1]
        out[f"{tag}_flat"] = pd.flatparam()       # [B,1,2A]
        out[f"{tag}_theta"] = np.concatenate([np.asarray(p).ravel() for p in para])
        out[f"{tag}_shapes"] = np.array([list(np.shape(p)) + [0] * (2 - np.ndim(p)) for p in para])
    np.savez_compressed(os.path.join(OUT, "model_fc_ref.npz"), **out)
    print("model_fc_ref.npz: param shapes", out["pend_shapes"].tolist())


# ---------------------------------------------------------------- PPO loss
def make_ppo_loss(tf, U, D):
    src = ref_lines("PPO.py", 29, 40)
    rng = np.random.default_rng(5)
    out = {}
    for tag, A, B, l_mul in (("a", 1, 64, 0.7), ("b", 6, 96, 1.0), ("c", 17, 50, 0.35)):
        flat = rng.standard_normal((B, 1, 2 * A)) * 0.5
        flat_old = flat + 0.15 * rng.standard_normal((B, 1, 2 * A))
        a = rng.standard_normal((B, 1, A))
        adv = rng.standard_normal((B, 1))
        etr = rng.standard_normal((B, 1))
        v = rng.standard_normal((B, 1)); v_old = v + 0.3 * rng.standard_normal((B, 1))
        res = {}
        for mode in ("canonical", "literal"):
            vv, vo = (v, v_old) if mode == "canonical" else (v[:, :, None], v_old[:, :, None])
            self = types.SimpleNamespace(a=a, adv=adv, etr=etr, pipredv=vv)
            ns = dict(tf=tf, U=U, np=np, self=self, piPD=D.DiagGaussianPd(flat),
                      oldpiPD=D.DiagGaussianPd(flat_old), oldpredv=vo,
                      CLIP_PARAM=0.2 * l_mul, ENTCOEFF=0.01, VCOEFF=0.5)   # PPO.py:19, main.py:15-17
            exec(src, ns)
            res[mode] = np.array([self.policyLoss, self.valueLoss, self.entropyLoss,
                                  self.total_loss], np.float64)
            if mode == "canonical":
                out[f"{tag}_ratio"] = ns["ratio"]
        out[f"{tag}_flat"], out[f"{tag}_flat_old"] = flat[:, 0], flat_old[:, 0]
        out[f"{tag}_act"], out[f"{tag}_adv"], out[f"{tag}_etr"] = a[:, 0], adv[:, 0], etr[:, 0]
        out[f"{tag}_v"], out[f"{tag}_v_old"] = v[:, 0], v_old[:, 0]
        out[f"{tag}_l_mul"] = np.array(l_mul)
        out[f"{tag}_losses"] = res["canonical"]            # policy, value, entropy, total
        out[f"{tag}_losses_literal"] = res["literal"]
    np.savez_compressed(os.path.join(OUT, "ppo_loss_ref.npz"), **out)
    print("ppo_loss_ref.npz: losses(a) canonical", out["a_losses"], "literal", out["a_losses_literal"])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    make_gae()
    tf, U, D, M, spaces = load_reference()
    make_distgauss(D)
    make_model_fc(tf, M, spaces)
    make_ppo_loss(tf, U, D)
    make_categorical(tf, U, D)
