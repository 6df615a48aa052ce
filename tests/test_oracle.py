"""The oracle restatement against (i) golden vectors produced by executing the reference's own
source (oracle/make_golden.py) and (ii) torch autograd for what the reference does not pin."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import dppo_oracle as O


def _npz(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# ------------------------------------------------------------------ GAE (Worker.py:82-92 verbatim)
def test_gae_matches_reference_lines(golden_dir):
    g = _npz(golden_dir, "gae_ref.npz")
    for i in range(int(g["n_cases"])):
        r, pv, d = g[f"c{i}_r"], g[f"c{i}_predv"], g[f"c{i}_done"]
        with np.errstate(all="ignore"):
            adv, etr = O.gae_rollout(r, pv, d, float(g["gamma"]), float(g["lam"]))
        np.testing.assert_array_equal(adv, g[f"c{i}_adv"], err_msg=f"case {i} adv")       # bit exact
        np.testing.assert_array_equal(etr, g[f"c{i}_etr"].astype(np.float32), err_msg=f"case {i} etr")
        # the same lines under NumPy-2 (NEP 50) promotion keep part of the scan in fp32: <= 2 ulp away
        ok = np.isfinite(g[f"c{i}_adv_np2"])
        np.testing.assert_allclose(adv[ok], g[f"c{i}_adv_np2"][ok], rtol=3e-6, atol=3e-6)
        np.testing.assert_allclose(etr, g[f"c{i}_etr_np2"], rtol=3e-6, atol=3e-6)


def test_gae_vectorized_equals_loop(golden_dir):
    g = _npz(golden_dir, "gae_ref.npz")
    cols = [i for i in range(int(g["n_cases"])) if len(g[f"c{i}_r"]) == 200]
    rew = np.stack([g[f"c{i}_r"] for i in cols], 1).astype(np.float32)
    pv = np.stack([g[f"c{i}_predv"] for i in cols], 1)
    dn = np.stack([g[f"c{i}_done"] for i in cols], 1)
    a1, e1 = O.gae_batch(rew, pv, dn, 0.99, 0.95)
    a2, e2 = O.gae_batch_vectorized(rew, pv, dn, 0.99, 0.95)
    np.testing.assert_allclose(a1, a2, rtol=2e-6, atol=2e-6)
    np.testing.assert_array_equal(e1, e2)


def test_gae_survey_kat():
    """REFERENCE-mode and STANDARD-mode raw advantages."""
    r = [1, .5, -1, 2, 0, 1.5]; pv = [.1, -.2, .3, .4, -.5, .6]; d = [0, 0, 1, 0, 0, 0]
    adv, etr = O.gae_rollout(r, pv, d, .99, .95, normalize=False)
    np.testing.assert_allclose(adv, [1.36035, 0.7, 2.3458133, 3.4554102, 2.499107, 1.494], rtol=1e-6)
    adv_s, _ = O.gae_rollout(r, pv, d, .99, .95, mode=O.GAE_STANDARD, last_v=.6, last_done=0,
                             normalize=False)
    np.testing.assert_allclose(adv_s, [0.48977618, -0.22565, -1.3, 3.45541013, 2.499107, 1.494], rtol=1e-6)


# ------------------------------------------------------------------ DiagGaussianPd
def test_diag_gaussian_matches_reference_class(golden_dir):
    g = _npz(golden_dir, "distgauss_ref.npz")
    p, x, q = g["fix_p"][None], g["fix_x"][None], g["fix_q"][None]
    np.testing.assert_allclose(O.dg_entropy(p), g["fix_entropy"], rtol=1e-14)
    np.testing.assert_allclose(O.dg_neglogp(p, x), g["fix_neglogp"], rtol=1e-14)
    np.testing.assert_allclose(O.dg_kl(p, q), g["fix_kl"], rtol=1e-13)
    # SURVEY Appendix C closed-form values
    assert abs(float(O.dg_entropy(p)[0]) - 6.175754132818691) < 1e-12
    assert abs(float(O.dg_neglogp(p, x)[0]) - 4.6382439888178055) < 1e-12
    assert abs(float(O.dg_kl(p, q)[0]) - 0.03006677708085259) < 1e-12
    for A in (1, 6, 17):
        fp, fq, xs = g[f"A{A}_p"], g[f"A{A}_q"], g[f"A{A}_x"]
        np.testing.assert_allclose(O.dg_neglogp(fp, xs), g[f"A{A}_neglogp"], rtol=1e-13)
        np.testing.assert_allclose(O.dg_logp(fp, xs), g[f"A{A}_logp"], rtol=1e-13)
        np.testing.assert_allclose(O.dg_entropy(fp), g[f"A{A}_entropy"], rtol=1e-13)
        np.testing.assert_allclose(O.dg_kl(fp, fq), g[f"A{A}_kl"], rtol=1e-12, atol=1e-15)
        np.testing.assert_array_equal(O.dg_mode(fp), g[f"A{A}_mode"])


def test_diag_gaussian_monte_carlo_identities():
    """distributions.py:270-295 (validate_probtype): entropy == -E[logp]; KL == -H - E_p[log q]."""
    np.random.seed(0)
    pdparam = np.array([-.2, .3, .4, -.5, .1, -.5, .1, 0.8])
    N = 100000
    M = np.repeat(pdparam[None], N, 0)
    X = O.dg_sample(M, np.random.randn(N, 4))
    ll = O.dg_logp(M, X)
    assert abs(O.dg_entropy(M).mean() + ll.mean()) < 3 * ll.std() / math.sqrt(N)
    q = pdparam + np.random.randn(pdparam.size) * 0.1
    M2 = np.repeat(q[None], N, 0)
    llq = O.dg_logp(M2, X)
    assert abs(O.dg_kl(M, M2).mean() - (-O.dg_entropy(M).mean() - llq.mean())) < 3 * llq.std() / math.sqrt(N)


# ------------------------------------------------------------------ Model.FC + normc
def test_mlp_forward_matches_reference_model_fc(golden_dir):
    g = _npz(golden_dir, "model_fc_ref.npz")
    for tag, obs_dim, A in (("pend", 3, 1), ("cheetah", 17, 6)):
        lay = O.make_layout(obs_dim, (16,), A)          # Model.py:12: one hidden layer of 16
        theta = g[f"{tag}_theta"]
        assert theta.size == lay.n_params
        shapes = [tuple(int(x) for x in row if x) for row in g[f"{tag}_shapes"]]
        assert shapes == [e[1] for e in lay.entries]    # creation order == flat layout order
        v, flat, _ = O.mlp_forward(theta, g[f"{tag}_s"], lay)
        np.testing.assert_allclose(v, g[f"{tag}_predv"][:, 0, 0], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(flat, g[f"{tag}_flat"][:, 0, :], rtol=1e-12, atol=1e-15)
        # normc_initializer(0.01): every kernel column has L2 norm 0.01 (tf_util.py:289)
        W1 = O.unflatten(theta, lay)["W1"]
        np.testing.assert_allclose(np.sqrt((W1 ** 2).sum(0)), 0.01, rtol=1e-5)
    assert O.make_layout(3, (16,), 1).n_params == 115
    assert O.make_layout(17, (256, 256), 6).n_params == 73741
    assert O.make_layout(376, (256, 256), 17).n_params == 171299
    assert O.make_layout(376, (1024,) * 4, 17).n_params == 3570723


def test_normc_columns_formula():
    np.random.seed(0)
    ref = np.random.randn(3, 16).astype(np.float32)
    ref *= 0.01 / np.sqrt(np.square(ref).sum(axis=0, keepdims=True))

    class R:                                         # same draw through the oracle's helper
        def standard_normal(self, shape):
            np.random.seed(0)
            return np.random.randn(*shape)
    np.testing.assert_array_equal(O.normc_columns((3, 16), 0.01, R()), ref)


# ------------------------------------------------------------------ PPO loss (PPO.py:29-40)
def test_ppo_loss_matches_reference_lines(golden_dir):
    g = _npz(golden_dir, "ppo_loss_ref.npz")
    for tag in "abc":
        losses, d_out, aux = O.ppo_losses_from_heads(
            g[f"{tag}_v"], g[f"{tag}_flat"], g[f"{tag}_v_old"], g[f"{tag}_flat_old"],
            g[f"{tag}_act"], g[f"{tag}_adv"], g[f"{tag}_etr"], float(g[f"{tag}_l_mul"]), O.HParams())
        np.testing.assert_allclose(losses, g[f"{tag}_losses"], rtol=1e-12)
        np.testing.assert_allclose(aux["ratio"], g[f"{tag}_ratio"][:, 0], rtol=1e-12)
        assert 0.02 < aux["clip_frac"] < 0.9
        # quirk D7: the literal pairwise value loss
        vl = O.value_loss_literal_pairwise(g[f"{tag}_v"], g[f"{tag}_v_old"], g[f"{tag}_etr"],
                                           float(g[f"{tag}_l_mul"]))
        np.testing.assert_allclose(vl, g[f"{tag}_losses_literal"][1], rtol=1e-12)


# ------------------------------------------------------------------ gradients: closed form vs autograd
def _torch_loss(theta, theta_old, obs, act, adv, ret, l_mul, lay, hp):
    def net(th):
        h = obs
        p = {}
        for name, shape, off in lay.entries:
            p[name] = th[off:off + int(np.prod(shape))].view(*shape)
        for l in range(len(lay.hidden)):
            h = torch.relu(h @ p[f"W{l + 1}"] + p[f"b{l + 1}"])
        return (h @ p["Wv"] + p["bv"])[:, 0], h @ p["Wp"] + p["bp"]

    def logp(flat):
        A = act.shape[1]
        return -(0.5 * (((act - flat[:, :A]) / torch.exp(flat[:, A:])) ** 2).sum(-1)
                 + 0.5 * math.log(2 * math.pi) * A + flat[:, A:].sum(-1))
    v, flat = net(theta)
    with torch.no_grad():
        vo, fo = net(theta_old)
    A = act.shape[1]
    c = hp.CLIP_PARAM * l_mul
    ent = (flat[:, A:] + 0.5 * math.log(2 * math.pi * math.e)).sum(-1)
    ratio = torch.exp(logp(flat) - logp(fo))
    pl = -torch.minimum(ratio * adv, torch.clamp(ratio, 1 - c, 1 + c) * adv).mean()
    el = -hp.ENTCOEFF * ent.mean()
    vl = hp.VCOEFF * torch.maximum((v - ret) ** 2, (vo + torch.clamp(v - vo, -c, c) - ret) ** 2).mean()
    return pl, vl, el, pl + el + vl


@pytest.mark.parametrize("obs_dim,hidden,A,B", [(3, (16,), 1, 200), (17, (32, 24), 6, 300),
                                                  (11, (8, 8, 8, 8), 3, 128), (5, (), 2, 64)])
def test_closed_form_grad_equals_autograd(obs_dim, hidden, A, B):
    lay = O.make_layout(obs_dim, hidden, A)
    th_pi, th_old = O.make_params_pair(lay, 1.0, 0.3, 0.05)
    rng = np.random.default_rng(1)
    obs = rng.standard_normal((B, obs_dim)); act = rng.standard_normal((B, A))
    adv = rng.standard_normal(B); ret = rng.standard_normal(B)
    hp = O.HParams()
    losses, grad, aux = O.ppo_loss_grad(th_pi, th_old, obs, act, adv, ret, 0.7, lay, hp)
    t = torch.tensor(th_pi, dtype=torch.float64, requires_grad=True)
    out = _torch_loss(t, torch.tensor(th_old, dtype=torch.float64), torch.tensor(obs),
                      torch.tensor(act), torch.tensor(adv), torch.tensor(ret), 0.7, lay, hp)
    out[3].backward()
    np.testing.assert_allclose(losses, [o.item() for o in out], rtol=1e-12)
    np.testing.assert_allclose(grad, t.grad.numpy(), rtol=1e-9, atol=1e-14)
    assert 0.02 < aux["clip_frac"] < 0.98 and aux["vclip_frac"] >= 0.0   # both branches live


def test_adam_tf_form_and_tower_mean():
    rng = np.random.default_rng(0)
    th = rng.standard_normal(50); m = np.zeros(50); v = np.zeros(50)
    gs = [rng.standard_normal(50) for _ in range(4)]
    g = O.tower_mean(gs)
    np.testing.assert_allclose(g, sum(gs) / 4)
    th1, m1, v1 = O.adam_tf(th, g, m, v, 1, 2e-5 * 0.7)
    # step 1 of TF Adam: lr_t = lr*sqrt(1-b2)/(1-b1); update = lr_t * m/(sqrt(v)+eps)
    lr_t = 2e-5 * 0.7 * math.sqrt(1 - 0.999) / (1 - 0.9)
    omb1 = float(np.float32(1) - np.float32(0.9)); omb2 = float(np.float32(1) - np.float32(0.999))   # TF: fp32 betas
    np.testing.assert_allclose(th1, th - lr_t * (omb1 * g) / (np.sqrt(omb2 * g * g) + 1e-8), rtol=1e-12)
    # differs from torch.optim.Adam's epsilon placement only at O(eps): sanity against torch
    tt = torch.tensor(th, requires_grad=True)
    opt = torch.optim.Adam([tt], lr=2e-5 * 0.7, eps=1e-8)
    tt.grad = torch.tensor(g); opt.step()
    np.testing.assert_allclose(th1, tt.detach().numpy(), rtol=0, atol=1e-7)


def test_lr_multiplier():
    assert O.lr_multiplier("constant", 10, 500) == 1.0
    assert O.lr_multiplier("linear", 125, 500) == 0.75
    assert O.lr_multiplier("linear", 900, 500) == 0


# ------------------------------------------------------------------ CategoricalPd (distributions.py:124-159) -- SURVEY 8 f4
def test_categorical_matches_reference_class(golden_dir):
    g = _npz(golden_dir, "categorical_ref.npz")
    p, q = g["fix_p"][None], g["fix_q"][None]
    np.testing.assert_allclose(O.cat_entropy(p), g["fix_entropy"], rtol=1e-13)
    np.testing.assert_allclose(O.cat_kl(p, q), g["fix_kl"], rtol=1e-10)
    assert O.cat_mode(p)[0] == g["fix_mode"][0] == 2
    for k in range(3):
        np.testing.assert_allclose(O.cat_neglogp(p, np.array([k])), g["fix_neglogp"][k], rtol=1e-13)
    for n in (2, 6, 18):
        fp, fq, x = g[f"n{n}_p"], g[f"n{n}_q"], g[f"n{n}_x"]
        np.testing.assert_allclose(O.cat_neglogp(fp, x), g[f"n{n}_neglogp"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(O.cat_logp(fp, x), g[f"n{n}_logp"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(O.cat_entropy(fp), g[f"n{n}_entropy"], rtol=1e-12)
        np.testing.assert_allclose(O.cat_kl(fp, fq), g[f"n{n}_kl"], rtol=1e-10, atol=1e-15)
        assert np.array_equal(O.cat_mode(fp), g[f"n{n}_mode"])


def test_categorical_monte_carlo_identities():
    """The reference's own self-check (distributions.py:260,269-295) on its categorical fixture, 3 sigma."""
    rng = np.random.default_rng(0)
    pdparam = np.array([-.2, .3, .5]); Nn = 100000
    M = np.repeat(pdparam[None], Nn, 0)
    X = O.cat_sample(M, rng.random(M.shape))
    ll = O.cat_logp(M, X)
    assert abs(O.cat_entropy(M).mean() + ll.mean()) < 3 * ll.std() / math.sqrt(Nn)
    q = pdparam + rng.standard_normal(3) * 0.1
    M2 = np.repeat(q[None], Nn, 0)
    ll2 = O.cat_logp(M2, X)
    assert abs(O.cat_kl(M, M2).mean() - (-O.cat_entropy(M).mean() - ll2.mean())) < 3 * ll2.std() / math.sqrt(Nn)


def test_categorical_ppo_loss_matches_reference_lines(golden_dir):
    g = _npz(golden_dir, "categorical_ref.npz")
    for tag in "abc":
        k = f"loss_{tag}_"
        losses, d_out, aux = O.ppo_losses_from_heads(g[k + "v"], g[k + "flat"], g[k + "v_old"], g[k + "flat_old"],
                                                     g[k + "act"], g[k + "adv"], g[k + "etr"], float(g[k + "l_mul"]),
                                                     O.HParams(), pd="cat")
        np.testing.assert_allclose(losses, g[k + "losses"], rtol=1e-12)
        np.testing.assert_allclose(aux["ratio"], g[k + "ratio"], rtol=1e-12)
        assert d_out.shape == (len(g[k + "v"]), 1 + g[k + "flat"].shape[1])


@pytest.mark.parametrize("obs_dim,hidden,ncat,B", [(4, (16,), 2, 200), (17, (32, 24), 6, 300), (5, (), 3, 64)])
def test_categorical_closed_form_grad_equals_autograd(obs_dim, hidden, ncat, B):
    lay = O.make_layout(obs_dim, hidden, ncat, pd="cat")
    assert lay.pd_width == ncat
    th_pi, th_old = O.make_params_pair(lay, 1.0, 0.5, 0.05)
    rng = np.random.default_rng(1)
    obs = rng.standard_normal((B, obs_dim)); act = rng.integers(0, ncat, (B, 1))
    adv = rng.standard_normal(B); ret = rng.standard_normal(B)
    hp = O.HParams()
    losses, grad, aux = O.ppo_loss_grad(th_pi, th_old, obs, act, adv, ret, 0.7, lay, hp)

    def net(th):
        h = torch.tensor(obs)
        p = {name: th[off:off + int(np.prod(shape))].view(*shape) for name, shape, off in lay.entries}
        for l in range(len(lay.hidden)):
            h = torch.relu(h @ p[f"W{l + 1}"] + p[f"b{l + 1}"])
        return (h @ p["Wv"] + p["bv"])[:, 0], h @ p["Wp"] + p["bp"]
    t = torch.tensor(th_pi, dtype=torch.float64, requires_grad=True)
    v, lg = net(t)
    with torch.no_grad():
        vo, lo = net(torch.tensor(th_old, dtype=torch.float64))
    a = torch.tensor(act[:, 0]); advt, rt = torch.tensor(adv), torch.tensor(ret)
    lsm, lsm_o = torch.log_softmax(lg, -1), torch.log_softmax(lo, -1)
    ratio = torch.exp(lsm.gather(1, a[:, None])[:, 0] - lsm_o.gather(1, a[:, None])[:, 0])
    c = hp.CLIP_PARAM * 0.7
    pl = -torch.minimum(ratio * advt, torch.clamp(ratio, 1 - c, 1 + c) * advt).mean()
    el = -hp.ENTCOEFF * (-(lsm.exp() * lsm).sum(-1)).mean()
    vl = hp.VCOEFF * torch.maximum((v - rt) ** 2, (vo + torch.clamp(v - vo, -c, c) - rt) ** 2).mean()
    (pl + el + vl).backward()
    np.testing.assert_allclose(losses, [pl.item(), vl.item(), el.item(), (pl + el + vl).item()], rtol=1e-12)
    np.testing.assert_allclose(grad, t.grad.numpy(), rtol=1e-9, atol=1e-14)
    assert 0.02 < aux["clip_frac"] < 0.98


# ------------------------------------------------------------------ epoch / minibatch scheduler restatement
def test_minibatch_perm_is_a_bijection_property():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(n=st.integers(1, 5000), seed=st.integers(0, 2**64 - 1), epoch=st.integers(0, 1000))
    def check(n, seed, epoch):
        p = O.minibatch_perm(n, seed, epoch)
        assert p.shape == (n,) and np.array_equal(np.sort(p), np.arange(n))
    check()
    a, b = O.minibatch_perm(4096, 1, 0), O.minibatch_perm(4096, 1, 1)
    assert (a != b).mean() > 0.99 and abs(np.corrcoef(a, np.arange(4096))[0, 1]) < 0.05


def test_update_epochs_whole_buffer_is_the_update_steps_loop():
    """n_epochs = UPDATE_STEPS, one whole-buffer minibatch, no shuffle == Chief.py:64: UPDATE_STEPS x [loss_grad, Adam]."""
    T, N, obs, hidden, A = 32, 4, 5, (8,), 2
    lay = O.make_layout(obs, hidden, A)
    th_pi, th_old = O.make_params_pair(lay)
    buf = O.make_buffer(T, N, obs, A)
    th, m, v, losses = O.update_epochs(th_pi, th_old, buf, 0.99, 0.95, 0.7, lay, 1e-3, n_epochs=3)
    adv, ret = O.gae_batch_vectorized(buf["rew"], buf["predv"], buf["done"], 0.99, 0.95)
    t2 = th_pi.astype(np.float64); m2 = np.zeros_like(t2); v2 = np.zeros_like(t2)
    for step in (1, 2, 3):
        l, g, _ = O.ppo_loss_grad(t2.astype(np.float32), th_old, buf["obs"], buf["act"], adv.reshape(-1), ret.reshape(-1), 0.7, lay)
        t2, m2, v2 = O.adam_tf(t2.astype(np.float32), g.astype(np.float32), m2, v2, step, 1e-3)
        np.testing.assert_allclose(losses[step - 1], l, rtol=1e-12)
    np.testing.assert_array_equal(th, t2)
    assert losses.shape == (3, 4) and not np.allclose(losses[0], losses[2])


def test_literal_pairwise_value_gradient_matches_autograd():
    """Quirk D7 as a trainable mode: the oracle's closed-form d(valueLoss)/dv of the pairwise broadcast (per rollout,
    tower mean over rollouts) against torch autograd of PPO.py:36-39 written with the literal ranks."""
    import torch
    rng = np.random.default_rng(5)
    B, G, c = 48, 4, 0.2 * 0.7
    v = rng.standard_normal(B); vo = v + 0.3 * rng.standard_normal(B); R = rng.standard_normal(B)
    loss, d_v, _ = O.value_literal_pairwise_grad(v, vo, R, 0.7, G)
    tv = torch.tensor(v, requires_grad=True); tvo = torch.tensor(vo); tR = torch.tensor(R)
    T = B // G
    tot = 0
    for g in range(G):
        idx = torch.arange(T) * G + g
        vi = tv[idx].reshape(T, 1, 1); voi = tvo[idx].reshape(T, 1, 1); etr = tR[idx].reshape(T, 1)   # [T,1,1] vs [T,1]
        vf1 = torch.square(vi - etr)                                                              # -> [T,T,1]
        vf2 = torch.square(voi + torch.clamp(vi - voi, -c, c) - etr)
        assert vf1.shape == (T, T, 1)
        tot = tot + 0.5 * torch.maximum(vf1, vf2).mean() / G
    tot.backward()
    assert abs(float(tot.detach()) - loss) < 1e-12
    np.testing.assert_allclose(tv.grad.numpy(), d_v, rtol=1e-10, atol=1e-14)
    assert abs(O.value_loss_literal_pairwise(v, vo, R, 0.7) - O.value_literal_pairwise_grad(v, vo, R, 0.7, 1)[0]) < 1e-14
