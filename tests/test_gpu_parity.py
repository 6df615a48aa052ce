"""GPU parity tests: every kernel of the update path, called through the C ABI, against the
oracle (and the committed reference-generated goldens).  fp32 tolerance from BASELINE.json
north_star: 1e-5 (relative to the tensor's max magnitude; stated per assert)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import dppo_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _engine(obs, hidden, act, max_rows, **kw):
    from dppo_b200 import DppoEngine
    return DppoEngine(obs, act, hidden, max_rows, **kw)


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).cuda()


def rel_err(got, ref):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-30))


def per_tensor_rel_err(got, ref, lay):
    worst = {}
    for name, shape, off in lay.entries:
        n = int(np.prod(shape))
        worst[name] = rel_err(got[off:off + n], ref[off:off + n])
    return worst


# ------------------------------------------------------------------------------- GAE
def test_gae_matches_reference_goldens(golden_dir):
    """Worker.py:82-92 executed verbatim (tests/golden/gae_ref.npz) vs dppo_gae, one rollout per call."""
    g = np.load(os.path.join(golden_dir, "gae_ref.npz"))
    eng = _engine(3, (16,), 1, 4096)
    for i in range(int(g["n_cases"])):
        r = g[f"c{i}_r"].astype(np.float32); pv = g[f"c{i}_predv"]; d = g[f"c{i}_done"]
        if not np.array_equal(r.astype(np.float64), g[f"c{i}_r"]):
            continue
        T = len(r)
        adv, ret = eng.gae(dev(r[:, None]), dev(pv[:, None]), dev(d[:, None]), 0.99, 0.95)
        adv = adv.cpu().numpy()[:, 0]; ret = ret.cpu().numpy()[:, 0]
        ref_adv = g[f"c{i}_adv"]
        np.testing.assert_allclose(ret, g[f"c{i}_etr"].astype(np.float32), rtol=0, atol=1e-6, err_msg=f"case {i}")
        ok = np.isfinite(ref_adv)
        if ok.all():
            np.testing.assert_allclose(adv, ref_adv, rtol=2e-6, atol=2e-6, err_msg=f"case {i} T={T}")
        else:           # degenerate std == 0 cases: the reference yields nan/inf, so must we
            assert not np.isfinite(adv[~ok]).any()


@pytest.mark.parametrize("T,N", [(200, 4), (2048, 64), (2048, 32), (1, 3), (17, 9), (1000, 13), (4100, 8)])
@pytest.mark.parametrize("mode", [0, 1])
def test_gae_batch_vs_oracle(T, N, mode):
    rng = np.random.default_rng(T * 1000 + N)
    rew = rng.standard_normal((T, N)).astype(np.float32)
    pv = rng.standard_normal((T, N)).astype(np.float32)
    done = (rng.random((T, N)) < 0.02).astype(np.float32)
    lv = rng.standard_normal(N).astype(np.float32); ld = (rng.random(N) < 0.3).astype(np.float32)
    eng = _engine(3, (16,), 1, max(T * N, 128))
    norm = T > 1
    adv, ret = eng.gae(dev(rew), dev(pv), dev(done), 0.99, 0.95, mode=mode, normalize=norm,
                       last_v=dev(lv), last_done=dev(ld))
    with np.errstate(all="ignore"):
        a_ref, r_ref = O.gae_batch(rew, pv, done, 0.99, 0.95, mode, lv, ld, normalize=norm)
    np.testing.assert_allclose(ret.cpu().numpy(), r_ref, rtol=0, atol=2e-6)
    np.testing.assert_allclose(adv.cpu().numpy(), a_ref, rtol=3e-6, atol=3e-6)
    # raw (un-normalised) advantages are bit-exact fp32 roundings of the fp64 scan
    adv_raw, _ = eng.gae(dev(rew), dev(pv), dev(done), 0.99, 0.95, mode=mode, normalize=False,
                         last_v=dev(lv), last_done=dev(ld))
    a_raw_ref, _ = O.gae_batch(rew, pv, done, 0.99, 0.95, mode, lv, ld, normalize=False)
    assert np.mean(adv_raw.cpu().numpy() != a_raw_ref) < 1e-3          # fma re-association: rare 1-ulp flips
    np.testing.assert_allclose(adv_raw.cpu().numpy(), a_raw_ref, rtol=2e-7, atol=1e-7)


# ------------------------------------------------------------------------------- GEMM building blocks
@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (1000, 256, 376), (777, 35, 256), (130, 3, 16), (64, 16, 3),
                                   (4096, 256, 256), (129, 257, 33)])
def test_gemm_modes(M, N, K):
    eng = _engine(376, (256, 256), 17, 4096)
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)); B = rng.standard_normal((K, N)); bias = rng.standard_normal(N)
    ref = np.maximum(A.astype(np.float32).astype(np.float64) @ B.astype(np.float32).astype(np.float64)
                     + bias.astype(np.float32), 0)
    got = eng.dbg_gemm(0, dev(A), dev(B), M, N, K, bias=dev(bias), relu=True).cpu().numpy()
    assert rel_err(got, ref) < 2e-6, "NN"
    # NT with mask:  C = (A . Bt^T) * [mask > 0]
    Bt = rng.standard_normal((N, K)); mask = rng.standard_normal((M, N))
    ref = (A.astype(np.float32).astype(np.float64) @ Bt.astype(np.float32).astype(np.float64).T) * (mask.astype(np.float32) > 0)
    got = eng.dbg_gemm(1, dev(A), dev(Bt), M, N, K, mask=dev(mask)).cpu().numpy()
    assert rel_err(got, ref) < 2e-6, "NT"
    # TN split-K:  C[M2,N] = X[Kr,M2]^T . G[Kr,N],  bias_out = colsum(G); needs M2*N+N <= n_params
    M2 = min(K, 300); Kr = M
    X = rng.standard_normal((Kr, M2)); G = rng.standard_normal((Kr, N))
    X64 = X.astype(np.float32).astype(np.float64); G64 = G.astype(np.float32).astype(np.float64)
    got, bo = eng.dbg_gemm(2, dev(X), dev(G), M2, N, Kr, want_bias_out=True)
    assert rel_err(got.cpu().numpy(), X64.T @ G64) < 2e-6, "TN"
    assert rel_err(bo.cpu().numpy(), G64.sum(0)) < 2e-6, "TN bias"


# ------------------------------------------------------------------------------- forward
def test_forward_matches_reference_model_fc_golden(golden_dir):
    """Model().FC executed from the reference source (tests/golden/model_fc_ref.npz)."""
    g = np.load(os.path.join(golden_dir, "model_fc_ref.npz"))
    for tag, obs_dim, A in (("pend", 3, 1), ("cheetah", 17, 6)):
        s = g[f"{tag}_s"].astype(np.float32)
        eng = _engine(obs_dim, (16,), A, 256)
        v, pd = eng.forward(dev(g[f"{tag}_theta"]), dev(s))
        lay = O.make_layout(obs_dim, (16,), A)
        v_ref, flat_ref, _ = O.mlp_forward(g[f"{tag}_theta"].astype(np.float32), s, lay)
        np.testing.assert_allclose(v.cpu().numpy(), v_ref, rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(pd.cpu().numpy(), flat_ref, rtol=1e-5, atol=1e-8)
        # and against the reference's own numbers (inputs there were fp64: looser)
        np.testing.assert_allclose(v.cpu().numpy(), g[f"{tag}_predv"][:, 0, 0], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("obs,hidden,A,B", [(3, (16,), 1, 800), (17, (256, 256), 6, 4096), (376, (256, 256), 17, 2048),
                                            (5, (), 2, 300), (33, (64, 48, 40, 24), 4, 1000),
                                            (376, (1024, 1024, 1024, 1024), 17, 512)])
def test_forward_vs_oracle(obs, hidden, A, B):
    lay = O.make_layout(obs, hidden, A)
    th, _ = O.make_params_pair(lay)
    x = np.random.default_rng(0).standard_normal((B, obs)).astype(np.float32)
    eng = _engine(obs, hidden, A, B)
    v, pd = eng.forward(dev(th), dev(x))
    v_ref, flat_ref, _ = O.mlp_forward(th, x, lay)
    assert rel_err(v.cpu().numpy(), v_ref) < TOL
    assert rel_err(pd.cpu().numpy(), flat_ref) < TOL


# ------------------------------------------------------------------------------- loss on reference goldens
def _identity_head_params(P, second_half):
    """L=0 net over obs = [head_pi | head_old] (2P wide) whose heads copy one half of the input."""
    A = (P - 1) // 2
    lay = O.make_layout(2 * P, (), A)
    th = np.zeros(lay.n_params, np.float32)
    p = O.unflatten(th, lay)
    base = P if second_half else 0
    p["Wv"][base + 0, 0] = 1.0
    for j in range(2 * A):
        p["Wp"][base + 1 + j, j] = 1.0
    return lay, th


def test_loss_matches_reference_ppo_lines_golden(golden_dir):
    """PPO.py:29-40 executed from the reference source (tests/golden/ppo_loss_ref.npz): feed its head
    values through identity heads so the device loss kernel sees exactly those numbers."""
    g = np.load(os.path.join(golden_dir, "ppo_loss_ref.npz"))
    for tag in "abc":
        flat, flat_old = g[f"{tag}_flat"], g[f"{tag}_flat_old"]
        B, A2 = flat.shape
        P = 1 + A2
        obs = np.concatenate([g[f"{tag}_v"][:, None], flat, g[f"{tag}_v_old"][:, None], flat_old], 1).astype(np.float32)
        lay, th_pi = _identity_head_params(P, False)
        _, th_old = _identity_head_params(P, True)
        eng = _engine(2 * P, (), A2 // 2, B)
        losses, _ = eng.loss_grad(dev(th_pi), dev(th_old), dev(obs), dev(g[f"{tag}_act"]), dev(g[f"{tag}_adv"]),
                                  dev(g[f"{tag}_etr"]), float(g[f"{tag}_l_mul"]))
        got = losses.cpu().numpy()
        np.testing.assert_allclose(got[:4], g[f"{tag}_losses"], rtol=2e-5, atol=2e-6)   # inputs rounded to fp32


def test_diag_gaussian_kl_entropy_golden(golden_dir):
    """DiagGaussianPd.kl / .entropy from the reference class (tests/golden/distgauss_ref.npz)."""
    g = np.load(os.path.join(golden_dir, "distgauss_ref.npz"))
    for A in (1, 6, 17):
        fp, fq, xs = g[f"A{A}_p"], g[f"A{A}_q"], g[f"A{A}_x"]
        B = fp.shape[0]; P = 1 + 2 * A
        z = np.zeros((B, 1))
        obs = np.concatenate([z, fq, z, fp], 1).astype(np.float32)      # pi = q, old = p  ->  KL(p || q)
        lay, th_pi = _identity_head_params(P, False)
        _, th_old = _identity_head_params(P, True)
        eng = _engine(2 * P, (), A, B)
        losses, _ = eng.loss_grad(dev(th_pi), dev(th_old), dev(obs), dev(xs), dev(np.zeros(B)), dev(np.zeros(B)), 1.0)
        got = losses.cpu().numpy()
        assert abs(got[4] - g[f"A{A}_kl"].mean()) < 2e-6 * max(1.0, abs(g[f"A{A}_kl"].mean()))
        ent_q = O.dg_entropy(fq.astype(np.float32).astype(np.float64)).mean()
        assert abs(got[2] - (-0.01 * ent_q)) < 1e-6


# ------------------------------------------------------------------------------- loss + gradient
CASES = [
    ("pendulum-cfg1", 200, 4, 3, (16,), 1),          # BASELINE configs[0]
    ("cheetah-slice", 256, 32, 17, (256, 256), 6),
    ("humanoid-slice", 128, 64, 376, (256, 256), 17),
    ("odd-shapes", 37, 5, 11, (24, 40, 8), 3),
    ("no-hidden", 64, 4, 5, (), 2),
    ("deep-1024-slice", 16, 64, 376, (1024, 1024, 1024, 1024), 17),
]


@pytest.mark.parametrize("name,T,N,obs,hidden,A", CASES)
def test_loss_grad_vs_oracle(name, T, N, obs, hidden, A):
    c = O.make_case(T, N, obs, hidden, A)
    lay = c["lay"]; B = T * N
    eng = _engine(obs, hidden, A, B)
    losses, g = eng.loss_grad(dev(c["theta_pi"]), dev(c["theta_old"]), dev(c["buf"]["obs"]), dev(c["act"]),
                              dev(c["adv"]), dev(c["ret"]), c["l_mul"])
    l_ref, g_ref, aux = O.ppo_loss_grad(c["theta_pi"], c["theta_old"], c["buf"]["obs"], c["act"], c["adv"], c["ret"],
                                        c["l_mul"], lay)
    got_l = losses.cpu().numpy(); got_g = g.cpu().numpy()[:lay.n_params]
    np.testing.assert_allclose(got_l[:4], l_ref, rtol=TOL, atol=TOL)
    assert abs(got_l[4] - aux["mean_kl"]) < 1e-5 * max(1, aux["mean_kl"])
    assert abs(got_l[5] - aux["clip_frac"]) < 1e-9 + 1.0 / B and abs(got_l[6] - aux["vclip_frac"]) < 1e-9 + 1.0 / B
    assert got_l[7] == B
    errs = per_tensor_rel_err(got_g, g_ref, lay)
    print(name, "per-tensor rel err", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs
    # branch coverage of the case itself
    assert aux["clip_frac"] > 0.01 and all(0.1 < float((h > 0).mean()) < 0.9 for h in aux["hs"])


@pytest.mark.parametrize("gemm_mode", [0, 1, 3])
def test_loss_grad_reference_regime_pi_equals_old(gemm_mode):
    """The reference's own regime (Appendix A-D9): pi == oldpi, normc(0.01) init -> ratio == 1, clips idle.  Activations are
    1e-2 ... 1e-5 here: the fp16-split path (mode 3) must hold the bar where `hi` alone would be subnormal."""
    lay = O.make_layout(17, (256, 256), 6)
    th = O.init_params(lay, np.random.default_rng(3))
    buf = O.make_buffer(64, 16, 17, 6, 5)
    adv, ret = O.gae_batch_vectorized(buf["rew"], buf["predv"], buf["done"], 0.99, 0.95)
    eng = _engine(17, (256, 256), 6, 1024, gemm_mode=gemm_mode)
    losses, g = eng.loss_grad(dev(th), dev(th), dev(buf["obs"]), dev(buf["act"]), dev(adv.reshape(-1)),
                              dev(ret.reshape(-1)), 1.0)
    l_ref, g_ref, aux = O.ppo_loss_grad(th, th, buf["obs"], buf["act"], adv.reshape(-1), ret.reshape(-1), 1.0, lay)
    got = losses.cpu().numpy()
    np.testing.assert_allclose(got[:4], l_ref, rtol=TOL, atol=TOL)
    assert got[5] == 0 and got[6] == 0 and abs(got[4]) < 1e-7
    assert max(per_tensor_rel_err(g.cpu().numpy()[:lay.n_params], g_ref, lay).values()) < TOL


def test_loss_grad_properties_full_size_humanoid():
    """BASELINE configs[2] at full size (T=2048, N=64, obs 376, act 17, 2x256) through
    size-independent properties: (i) the gradient of the whole batch equals the mean of the gradients
    of its two halves (tower mean, PPO.py:55-65), (ii) deterministic re-execution is bit-identical,
    (iii) a slice of rows agrees with the fp64 oracle."""
    T, N, obs, hidden, A = 2048, 64, 376, (256, 256), 17
    lay = O.make_layout(obs, hidden, A)
    th_pi, th_old = O.make_params_pair(lay)
    buf = O.make_buffer(T, N, obs, A)
    B = T * N
    eng = _engine(obs, hidden, A, B)
    d_rew, d_pv, d_done = dev(buf["rew"]), dev(buf["predv"]), dev(buf["done"])
    adv, ret = eng.gae(d_rew, d_pv, d_done, 0.99, 0.95)
    d_obs, d_act = dev(buf["obs"]), dev(buf["act"])
    d_pi, d_old = dev(th_pi), dev(th_old)
    adv = adv.reshape(-1); ret = ret.reshape(-1)
    _, g_all = eng.loss_grad(d_pi, d_old, d_obs, d_act, adv, ret, 0.7)
    g_all = g_all.clone()
    _, g_again = eng.loss_grad(d_pi, d_old, d_obs, d_act, adv, ret, 0.7)
    assert torch.equal(g_all, g_again), "update is not deterministic"
    h = B // 2
    _, g0 = eng.loss_grad(d_pi, d_old, d_obs[:h], d_act[:h], adv[:h], ret[:h], 0.7); g0 = g0.clone()
    _, g1 = eng.loss_grad(d_pi, d_old, d_obs[h:], d_act[h:], adv[h:], ret[h:], 0.7)
    mean = 0.5 * (g0.double() + g1.double()).cpu().numpy()
    n = lay.n_params
    assert max(per_tensor_rel_err(g_all.cpu().numpy()[:n], mean[:n], lay).values()) < TOL
    np.testing.assert_allclose(g_all.cpu().numpy()[n:n + 7], mean[n:n + 7], rtol=TOL, atol=TOL)
    # (iii) 8192-row slice against the oracle, with branch-ambiguous rows edited out
    s = slice(0, 8192)
    a_np = adv[s].cpu().numpy(); r_np = ret[s].cpu().numpy()
    obs_s, _ = O.sanitize_obs(th_pi, buf["obs"][s], lay)
    act_s, ret_s, _ = O.sanitize_rows(th_pi, th_old, obs_s, buf["act"][s], a_np, r_np, 0.7, lay)
    losses, g_s = eng.loss_grad(d_pi, d_old, dev(obs_s), dev(act_s), adv[s].contiguous(), dev(ret_s), 0.7)
    l_ref, g_ref, _ = O.ppo_loss_grad(th_pi, th_old, obs_s, act_s, a_np, ret_s, 0.7, lay)
    np.testing.assert_allclose(losses.cpu().numpy()[:4], l_ref, rtol=TOL, atol=TOL)
    assert max(per_tensor_rel_err(g_s.cpu().numpy()[:n], g_ref, lay).values()) < TOL


def test_full_size_cheetah_vs_oracle():
    """BASELINE configs[1] at full size: T=2048, N=32, obs 17, act 6, 2x256, B=65536 rows."""
    c = O.make_case(2048, 32, 17, (256, 256), 6)
    lay = c["lay"]
    eng = _engine(17, (256, 256), 6, 65536)
    adv_d, ret_d = eng.gae(dev(c["buf"]["rew"]), dev(c["buf"]["predv"]), dev(c["buf"]["done"]), 0.99, 0.95)
    np.testing.assert_allclose(adv_d.cpu().numpy().reshape(-1), c["adv"], rtol=3e-6, atol=3e-6)
    losses, g = eng.loss_grad(dev(c["theta_pi"]), dev(c["theta_old"]), dev(c["buf"]["obs"]), dev(c["act"]),
                              dev(c["adv"]), dev(c["ret"]), c["l_mul"])
    l_ref, g_ref, _ = O.ppo_loss_grad(c["theta_pi"], c["theta_old"], c["buf"]["obs"], c["act"], c["adv"], c["ret"],
                                      c["l_mul"], lay)
    np.testing.assert_allclose(losses.cpu().numpy()[:4], l_ref, rtol=TOL, atol=TOL)
    assert max(per_tensor_rel_err(g.cpu().numpy()[:lay.n_params], g_ref, lay).values()) < TOL


# ------------------------------------------------------------------------------- Adam / sync / allreduce / act
def test_adam_sync_allreduce_world1():
    lay = O.make_layout(17, (256, 256), 6)
    n = lay.n_params
    rng = np.random.default_rng(0)
    th = rng.standard_normal(n).astype(np.float32); g = (rng.standard_normal(n) * 1e-2).astype(np.float32)
    eng = _engine(17, (256, 256), 6, 256)
    p, m, v = dev(th), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    th_ref, m_ref, v_ref = th.astype(np.float64), np.zeros(n), np.zeros(n)
    for step in (1, 2, 3):
        eng.adam_apply(p, dev(g), m, v, step, 2e-5 * 0.7)
        th_ref, m_ref, v_ref = O.adam_tf(th_ref, g, m_ref, v_ref, step, 2e-5 * 0.7)
    # fp32 parameters: the update (4e-5) is resolved to one ulp of the O(1) parameter values
    np.testing.assert_allclose(p.cpu().numpy(), th_ref, rtol=1.2e-7, atol=1e-7)
    small = np.abs(th) < 0.25                         # where fp32 resolves the 3-step update itself to <1%
    upd = p.cpu().numpy().astype(np.float64) - th
    assert rel_err(upd[small], (th_ref - th)[small]) < 1e-2
    np.testing.assert_allclose(m.cpu().numpy(), m_ref, rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(v.cpu().numpy(), v_ref, rtol=1e-5, atol=1e-15)
    old = torch.zeros(n, device="cuda")
    eng.sync_old(p, old)
    assert torch.equal(old, p)
    # world = 1 communicator: average of one tower is the tower
    eng.comm_init(0, 1, eng.unique_id())
    buf = dev(g).clone()
    eng.allreduce_avg(buf)
    torch.cuda.synchronize()
    assert torch.equal(buf, dev(g))


def test_act_sampling_moments_and_monte_carlo_identity():
    """DiagGaussianPd.sample (distributions.py:204-205) + the reference self-check identity
    entropy == -E[logp] (distributions.py:279-283), N=100000, 3 sigma."""
    A = 4; P = 1 + 2 * A
    pdparam = np.array([-.2, .3, .4, -.5, .1, -.5, .1, 0.8])
    Nn = 100000
    lay, th = _identity_head_params(P, False)
    obs = np.zeros((Nn, 2 * P), np.float32); obs[:, 1:P] = pdparam
    eng = _engine(2 * P, (), A, Nn)
    a, v = eng.act(dev(th), dev(obs), seed=1234, offset=0)
    a2, _ = eng.act(dev(th), dev(obs), seed=1234, offset=0)
    a3, _ = eng.act(dev(th), dev(obs), seed=1234, offset=1)
    assert torch.equal(a, a2) and not torch.equal(a, a3)
    X = a.cpu().numpy().astype(np.float64)
    M = np.repeat(pdparam[None], Nn, 0)
    ll = O.dg_logp(M, X)
    assert abs(O.dg_entropy(M).mean() + ll.mean()) < 3 * ll.std() / math.sqrt(Nn)
    np.testing.assert_allclose(X.mean(0), pdparam[:A], atol=4 * np.exp(pdparam[A:]).max() / math.sqrt(Nn))
    np.testing.assert_allclose(X.std(0), np.exp(pdparam[A:]), rtol=0.02)
    assert float(v.abs().max()) == 0.0


def test_update_host_equals_device_path():
    """dppo_update_host (pinned host rollout, chunked H2D overlapped with compute) must produce
    the same losses/gradient as gae + loss_grad on device-resident buffers."""
    T, N, obs, hidden, A = 512, 64, 376, (256, 256), 17
    lay = O.make_layout(obs, hidden, A)
    th_pi, th_old = O.make_params_pair(lay)
    buf = O.make_buffer(T, N, obs, A)
    eng = _engine(obs, hidden, A, T * N)
    d_pi, d_old = dev(th_pi), dev(th_old)
    adv, ret = eng.gae(dev(buf["rew"]), dev(buf["predv"]), dev(buf["done"]), 0.99, 0.95)
    l_dev, g_dev = eng.loss_grad(d_pi, d_old, dev(buf["obs"]), dev(buf["act"]), adv.reshape(-1), ret.reshape(-1), 0.7)
    g_dev = g_dev.clone()
    pin = {k: torch.from_numpy(v).pin_memory() for k, v in buf.items()}
    l_h, g_h = eng.update_host(d_pi, d_old, pin["obs"], pin["act"], pin["rew"], pin["predv"], pin["done"],
                               0.99, 0.95, 0.7)
    n = lay.n_params
    assert max(per_tensor_rel_err(g_h.cpu().numpy()[:n], g_dev.cpu().numpy()[:n], lay).values()) < TOL
    np.testing.assert_allclose(np.array(l_h), g_dev.cpu().numpy()[n:], rtol=1e-6, atol=1e-7)
    # with the optimiser inside the call: same parameters as adam_apply on the returned gradient
    p1 = d_pi.clone(); m1 = torch.zeros(n, device="cuda"); v1 = torch.zeros(n, device="cuda")
    eng.adam_apply(p1, g_h, m1, v1, 1, 1e-4)
    p2 = d_pi.clone(); m2 = torch.zeros(n, device="cuda"); v2 = torch.zeros(n, device="cuda")
    _, g2 = eng.update_host(p2, d_old, pin["obs"], pin["act"], pin["rew"], pin["predv"], pin["done"], 0.99, 0.95, 0.7,
                            adam_m=m2, adam_v=v2, step=1, lr=1e-4)
    assert torch.equal(g2[:n], g_h[:n]) and torch.equal(p1, p2) and torch.equal(m1, m2) and torch.equal(v1, v2)


def test_errors_are_loud():
    eng = _engine(3, (16,), 1, 128)
    lay = O.make_layout(3, (16,), 1)
    th = dev(np.zeros(lay.n_params))
    with pytest.raises(RuntimeError, match="max_rows"):
        eng.forward(th, dev(np.zeros((256, 3))))
    with pytest.raises((TypeError, ValueError)):
        eng.forward(th.cpu(), dev(np.zeros((8, 3))))
    with pytest.raises(RuntimeError, match="communicator"):
        eng.allreduce_avg(th)
    with pytest.raises(RuntimeError, match="STANDARD"):
        z = dev(np.zeros((4, 2)))
        eng.gae(z, z, z, 0.99, 0.95, mode=1)


def test_diag_gaussian_methods_golden(golden_dir):
    """dppo_diag_gaussian (the Pd methods as stand-alone ops) vs the reference class' own outputs."""
    g = np.load(os.path.join(golden_dir, "distgauss_ref.npz"))
    eng = _engine(3, (16,), 1, 128)
    for A in (1, 6, 17):
        fp, fq, xs = dev(g[f"A{A}_p"]), dev(g[f"A{A}_q"]), dev(g[f"A{A}_x"])
        np.testing.assert_allclose(eng.diag_gaussian("neglogp", fp, xs).cpu().numpy(), g[f"A{A}_neglogp"], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(eng.diag_gaussian("entropy", fp).cpu().numpy(), g[f"A{A}_entropy"], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(eng.diag_gaussian("kl", fp, fq).cpu().numpy(), g[f"A{A}_kl"], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(eng.diag_gaussian("mode", fp).cpu().numpy(), g[f"A{A}_mode"].astype(np.float32))
    # reference self-check fixture, SURVEY Appendix C values
    p = dev(g["fix_p"][None]); x = dev(g["fix_x"][None]); q = dev(g["fix_q"][None])
    assert abs(float(eng.diag_gaussian("entropy", p)[0]) - 6.175754132818691) < 1e-5
    assert abs(float(eng.diag_gaussian("neglogp", p, x)[0]) - 4.6382439888178055) < 1e-5
    assert abs(float(eng.diag_gaussian("kl", p, q)[0]) - 0.03006677708085259) < 1e-6
    # KL identity of validate_probtype (distributions.py:286-295) with device samples
    Nn = 100000
    M = dev(np.repeat(g["fix_p"][None], Nn, 0)); M2 = dev(np.repeat(g["fix_q"][None], Nn, 0))
    X = eng.diag_gaussian("sample", M, seed=7, offset=3)
    llq = -eng.diag_gaussian("neglogp", M2, X).double().cpu().numpy()
    ent = eng.diag_gaussian("entropy", M).double().cpu().numpy().mean()
    kl = eng.diag_gaussian("kl", M, M2).double().cpu().numpy().mean()
    assert abs(kl - (-ent - llq.mean())) < 3 * llq.std() / math.sqrt(Nn)


# ------------------------------------------------------------------------------- multi-GPU (NCCL over NVLink)
def _multi_gpu_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from dppo_b200 import DppoEngine
    from dppo_b200.dist import init_engine_comm, shard_rollouts
    T, N, obs, hidden, A = 128, 16, 17, (64, 64), 6
    c = O.make_case(T, N, obs, hidden, A)
    lay = c["lay"]
    lo, hi = shard_rollouts(N, rank, world)
    rows = (np.arange(T)[:, None] * N + np.arange(lo, hi)[None, :]).reshape(-1)
    eng = DppoEngine(obs, A, hidden, len(rows), device=rank)
    init_engine_comm(eng, rank, world)
    d = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).cuda(rank)   # noqa: E731
    p_pi = d(c["theta_pi"])
    _, g = eng.loss_grad(p_pi, d(c["theta_old"]), d(c["buf"]["obs"][rows]), d(c["act"][rows]), d(c["adv"][rows]),
                         d(c["ret"][rows]), c["l_mul"])
    eng.allreduce_avg(g)                                           # PPO._average_gradients across ranks
    m = torch.zeros(lay.n_params, device=f"cuda:{rank}"); v = torch.zeros_like(m)
    eng.adam_apply(p_pi, g, m, v, 1, 2e-5 * c["l_mul"])
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"g_{rank}.npy"), g.cpu().numpy())
    np.save(os.path.join(out_dir, f"p_{rank}.npy"), p_pi.cpu().numpy())
    # whole update in one call: NVLink peer-memory allreduce fused with Adam (3 consecutive updates exercise the
    # double-buffered exchange) vs the same update through ncclAllReduce + adam_kernel
    n = lay.n_params
    rb = {k: d(c["buf"][k][:, lo:hi]) for k in ("rew", "predv", "done")}
    obs_r, act_r, old_r = d(c["buf"]["obs"][rows]), d(c["act"][rows]), d(c["theta_old"])
    pa = d(c["theta_pi"]); ma = torch.zeros(n, device=f"cuda:{rank}"); va = torch.zeros_like(ma)
    pb = d(c["theta_pi"]); mb = torch.zeros_like(ma); vb = torch.zeros_like(ma)
    for step in (1, 2, 3):
        _, ga = eng.update(pa, old_r, obs_r, act_r, rb["rew"], rb["predv"], rb["done"], 0.99, 0.95, 0.7,
                           adam_m=ma, adam_v=va, step=step, lr=1e-3)
        ga = ga.clone()
        adv_b, ret_b = eng.gae(rb["rew"], rb["predv"], rb["done"], 0.99, 0.95)
        _, gb = eng.loss_grad(pb, old_r, obs_r, act_r, adv_b.reshape(-1), ret_b.reshape(-1), 0.7)
        eng.allreduce_avg(gb)
        eng.adam_apply(pb, gb, mb, vb, step, 1e-3)
        torch.cuda.synchronize()
        np.testing.assert_allclose(ga.cpu().numpy(), gb.cpu().numpy(), rtol=2e-6, atol=2e-8)   # sum*1/W vs ncclAvg rounding
        np.testing.assert_allclose(pa.cpu().numpy(), pb.cpu().numpy(), rtol=0, atol=1e-6)
    np.save(os.path.join(out_dir, f"pa_{rank}.npy"), pa.cpu().numpy())
    # the device scheduler across ranks: 2 epochs x shuffled minibatches, one exchange per minibatch, == three more
    # single updates when the minibatch is the whole shard and nothing is shuffled
    pc = pa.clone(); mc = ma.clone(); vc = va.clone()
    for step in (4, 5):
        eng.update(pa, old_r, obs_r, act_r, rb["rew"], rb["predv"], rb["done"], 0.99, 0.95, 0.7, adam_m=ma, adam_v=va,
                   step=step, lr=1e-3)
    eng.update_epochs(pc, old_r, obs_r, act_r, rb["rew"], rb["predv"], rb["done"], 0.99, 0.95, 0.7, mc, vc, step0=4, lr=1e-3,
                      n_epochs=2)
    torch.cuda.synchronize()
    assert torch.equal(pa, pc) and torch.equal(ma, mc)
    le, _ = eng.update_epochs(pc, old_r, obs_r, act_r, rb["rew"], rb["predv"], rb["done"], 0.99, 0.95, 0.7, mc, vc, step0=6,
                              lr=1e-3, n_epochs=2, minibatch_rows=max(len(rows) // 4, 1), shuffle=True, seed=11)
    torch.cuda.synchronize()
    assert le.shape[0] == 8 and torch.isfinite(le).all() and torch.isfinite(pc).all()
    np.save(os.path.join(out_dir, f"pc_{rank}.npy"), pc.cpu().numpy())
    dist.destroy_process_group()


def test_multi_gpu_allreduce_equals_tower_mean(tmp_path):
    """N ranks, one tower each: the NCCL-averaged gradient equals the oracle's tower mean (PPO.py:55-65)
    and every rank ends with bit-identical parameters (no chief->worker broadcast needed)."""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_multi_gpu_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    T, N, obs, hidden, A = 128, 16, 17, (64, 64), 6
    c = O.make_case(T, N, obs, hidden, A)
    lay = c["lay"]
    from dppo_b200.dist import shard_rollouts
    towers = []
    for r in range(world):
        lo, hi = shard_rollouts(N, r, world)
        rows = (np.arange(T)[:, None] * N + np.arange(lo, hi)[None, :]).reshape(-1)
        losses, g, _ = O.ppo_loss_grad(c["theta_pi"], c["theta_old"], c["buf"]["obs"][rows], c["act"][rows],
                                       c["adv"][rows], c["ret"][rows], c["l_mul"], lay)
        towers.append(np.concatenate([g, losses]))
    ref = O.tower_mean(towers)
    gs = [np.load(tmp_path / f"g_{r}.npy") for r in range(world)]
    ps = [np.load(tmp_path / f"p_{r}.npy") for r in range(world)]
    pas = [np.load(tmp_path / f"pa_{r}.npy") for r in range(world)]
    for r in range(1, world):
        np.testing.assert_array_equal(gs[0], gs[r]); np.testing.assert_array_equal(ps[0], ps[r])
        np.testing.assert_array_equal(pas[0], pas[r])          # peer-memory path: bit-identical parameters on all ranks
        np.testing.assert_array_equal(np.load(tmp_path / "pc_0.npy"), np.load(tmp_path / f"pc_{r}.npy"))   # after the scheduler
    n = lay.n_params
    assert max(per_tensor_rel_err(gs[0][:n], ref[:n], lay).values()) < TOL
    np.testing.assert_allclose(gs[0][n:n + 4], ref[n:], rtol=TOL, atol=TOL)


# ------------------------------------------------------------------------------- tensor-core (tcgen05) products
@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (256, 256, 64), (1000, 256, 376), (777, 35, 256), (130, 128, 100),
                                   (4096, 256, 256), (129, 200, 33), (300, 64, 35)])
@pytest.mark.parametrize("terms", [1, 3])
def test_tc_gemm_k_major(M, N, K, terms):
    """tcgen05 kind::tf32 products with both operands K-contiguous (forward / dX); terms=1 checks the smem
    layouts and descriptors at TF32 accuracy, terms=3 is the fp32-parity split."""
    eng = _engine(376, (256, 256), 17, 4096)
    base = 16 if terms == 3 else 32
    tol = 6e-6 if terms == 3 else 2e-3   # TMEM accumulation truncates: ~K/8 biased half-ulps
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K)).astype(np.float32); Bt = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32); mask = rng.standard_normal((M, N)).astype(np.float32)
    ref = A.astype(np.float64) @ Bt.astype(np.float64).T
    got = eng.dbg_gemm(base + 0, dev(A), dev(Bt), M, N, K, bias=dev(bias), relu=True).cpu().numpy()
    assert rel_err(got, np.maximum(ref + bias, 0)) < tol, "kk/NN epilogue"
    got = eng.dbg_gemm(base + 1, dev(A), dev(Bt), M, N, K, mask=dev(mask)).cpu().numpy()
    assert rel_err(got, ref * (mask > 0)) < tol, "kk/NT epilogue"


@pytest.mark.parametrize("M,N,Kr", [(128, 256, 32), (376, 256, 1000), (256, 35, 2048), (120, 64, 333), (256, 256, 4096)])
@pytest.mark.parametrize("terms", [1, 3])
def test_tc_gemm_mn_major(M, N, Kr, terms):
    """dW product C[M,N] = X[Kr,M]^T . G[Kr,N] with both operands MN-contiguous (SWIZZLE_128B_BASE32B), split-K."""
    eng = _engine(376, (256, 256), 17, 4096)
    base = 16 if terms == 3 else 32
    tol = 6e-6 if terms == 3 else 2e-3   # TMEM accumulation truncates: ~K/8 biased half-ulps
    rng = np.random.default_rng(M + N + Kr)
    X = rng.standard_normal((Kr, M)).astype(np.float32); G = rng.standard_normal((Kr, N)).astype(np.float32)
    got, bo = eng.dbg_gemm(base + 2, dev(X), dev(G), M, N, Kr, want_bias_out=True)
    assert rel_err(got.cpu().numpy(), X.astype(np.float64).T @ G.astype(np.float64)) < tol
    assert rel_err(bo.cpu().numpy(), G.astype(np.float64).sum(0)) < 2e-6


@pytest.mark.parametrize("name,T,N,obs,hidden,A", CASES)
def test_loss_grad_vs_oracle_tensor_core(name, T, N, obs, hidden, A):
    """The whole tower on the tcgen05 3xTF32 path (DPPO_GEMM_TF32X3) against the fp64 oracle, same 1e-5 bar."""
    c = O.make_case(T, N, obs, hidden, A)
    lay = c["lay"]; B = T * N
    eng = _engine(obs, hidden, A, B, gemm_mode=1)
    losses, g = eng.loss_grad(dev(c["theta_pi"]), dev(c["theta_old"]), dev(c["buf"]["obs"]), dev(c["act"]),
                              dev(c["adv"]), dev(c["ret"]), c["l_mul"])
    l_ref, g_ref, aux = O.ppo_loss_grad(c["theta_pi"], c["theta_old"], c["buf"]["obs"], c["act"], c["adv"], c["ret"],
                                        c["l_mul"], lay)
    np.testing.assert_allclose(losses.cpu().numpy()[:4], l_ref, rtol=TOL, atol=TOL)
    errs = per_tensor_rel_err(g.cpu().numpy()[:lay.n_params], g_ref, lay)
    print(name, "tc per-tensor rel err", {k: f"{v:.2e}" for k, v in errs.items()})
    # 1e-5 holds for the fp32 configs (H <= 256); the 4x1024 net (BASELINE configs[3], the bf16 config) reduces over
    # K=1024 per layer: 128 truncating TMEM adds per accumulator -> 1.5e-5 measured, bound stated here
    assert max(errs.values()) < (TOL if max(hidden, default=0) <= 256 else 3e-5), errs


@pytest.mark.parametrize("M,N,Kr", [(128, 256, 32), (376, 256, 1000), (256, 36, 2048), (120, 64, 333), (256, 256, 4096),
                                    (376, 256, 16), (64, 128, 7)])
def test_tc_gemm_tn_tma(M, N, Kr):
    """TMA-fed dW product (cp.async.bulk.tensor -> SWIZZLE_128B_ATOM_32B tiles, in-place hi/lo split, fused colsum)."""
    eng = _engine(376, (256, 256), 17, 4096)
    rng = np.random.default_rng(M + N + Kr)
    X = rng.standard_normal((Kr, M)).astype(np.float32); G = rng.standard_normal((Kr, N)).astype(np.float32)
    got, bo = eng.dbg_gemm(19, dev(X), dev(G), M, N, Kr, want_bias_out=True)
    assert rel_err(got.cpu().numpy(), X.astype(np.float64).T @ G.astype(np.float64)) < 6e-6
    assert rel_err(bo.cpu().numpy(), G.astype(np.float64).sum(0)) < 2e-6


@pytest.mark.parametrize("gemm_mode", [0, 1])
def test_update_single_call_equals_separate_calls(gemm_mode):
    """dppo_update (GAE on a side stream under the forward GEMMs, then loss/backward/Adam) == the five separate calls."""
    T, N, obs, hidden, A = 256, 64, 376, (256, 256), 17
    lay = O.make_layout(obs, hidden, A)
    th_pi, th_old = O.make_params_pair(lay)
    buf = O.make_buffer(T, N, obs, A)
    eng = _engine(obs, hidden, A, T * N, gemm_mode=gemm_mode)
    d = {k: dev(v) for k, v in buf.items()}
    n = lay.n_params
    # separate calls
    p1 = dev(th_pi); m1 = torch.zeros(n, device="cuda"); v1 = torch.zeros(n, device="cuda")
    adv, ret = eng.gae(d["rew"], d["predv"], d["done"], 0.99, 0.95)
    _, g1 = eng.loss_grad(p1, dev(th_old), d["obs"], d["act"], adv.reshape(-1), ret.reshape(-1), 0.7)
    g1 = g1.clone()
    eng.adam_apply(p1, g1, m1, v1, 1, 2e-5 * 0.7)
    # one call, twice (the second exercises the stream hand-over between consecutive updates)
    p2 = dev(th_pi); m2 = torch.zeros(n, device="cuda"); v2 = torch.zeros(n, device="cuda")
    _, g2 = eng.update(p2, dev(th_old), d["obs"], d["act"], d["rew"], d["predv"], d["done"], 0.99, 0.95, 0.7,
                       adam_m=m2, adam_v=v2, step=1, lr=2e-5 * 0.7)
    g2 = g2.clone()
    torch.cuda.synchronize()
    assert torch.equal(g1, g2) and torch.equal(p1, p2) and torch.equal(m1, m2) and torch.equal(v1, v2)
    eng.update(p2, dev(th_old), d["obs"], d["act"], d["rew"], d["predv"], d["done"], 0.99, 0.95, 0.7,
               adam_m=m2, adam_v=v2, step=2, lr=2e-5 * 0.7)
    torch.cuda.synchronize()
    assert torch.isfinite(p2).all() and not torch.equal(p1, p2)


# ------------------------------------------------------------------------------- bf16 tensor-core path (BASELINE configs[3])
def _bf16(x):
    return torch.as_tensor(x, dtype=torch.float32).to(torch.bfloat16).to(torch.float64).numpy()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1000, 256, 376), (777, 35, 256), (4096, 1024, 1024), (129, 200, 33)])
def test_bf16_gemm_k_major(M, N, K):
    """kind::f16 bf16 products, K-major operands: exact vs an fp64 product of the bf16-rounded operands (fp32 accumulate)."""
    eng = _engine(376, (256, 256), 17, 4096)
    rng = np.random.default_rng(M + 3 * N + K)
    A = rng.standard_normal((M, K)).astype(np.float32); Bt = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32); mask = rng.standard_normal((M, N)).astype(np.float32)
    ref = _bf16(A) @ _bf16(Bt).T
    got = eng.dbg_gemm(64, dev(A), dev(Bt), M, N, K, bias=dev(bias), relu=True).cpu().numpy()
    assert rel_err(got, np.maximum(ref + bias, 0)) < 3e-6
    got = eng.dbg_gemm(65, dev(A), dev(Bt), M, N, K, mask=dev(mask)).cpu().numpy()
    assert rel_err(got, ref * (mask > 0)) < 3e-6


@pytest.mark.parametrize("M,N,Kr", [(128, 256, 32), (376, 256, 1000), (256, 35, 2048), (120, 64, 333), (1024, 1024, 512)])
def test_bf16_gemm_mn_major(M, N, Kr):
    eng = _engine(376, (1024, 1024), 17, 4096)
    rng = np.random.default_rng(M + N + Kr)
    X = rng.standard_normal((Kr, M)).astype(np.float32); G = rng.standard_normal((Kr, N)).astype(np.float32)
    got = eng.dbg_gemm(66, dev(X), dev(G), M, N, Kr).cpu().numpy()
    assert rel_err(got, _bf16(X).T @ _bf16(G)) < 3e-6


@pytest.mark.parametrize("name,T,N,obs,hidden,A", [CASES[0], CASES[2], CASES[5]])
def test_loss_grad_bf16_path(name, T, N, obs, hidden, A):
    """DPPO_GEMM_BF16 tower vs the oracle run with the same operand rounding (every matmul operand rounded to bf16,
    wide accumulation).  The few rows whose clip / ReLU branch sits within bf16-accumulation noise of a switch may
    flip, so the bound is 2e-3 of each tensor's max magnitude; the distance to the un-rounded fp64 oracle (inherent to
    8-bit mantissas: the PPO ratio is an exp of a difference of two rounded nets) is printed, not asserted."""
    c = O.make_case(T, N, obs, hidden, A)
    lay = c["lay"]; B = T * N
    eng = _engine(obs, hidden, A, B, gemm_mode=2)
    losses, g = eng.loss_grad(dev(c["theta_pi"]), dev(c["theta_old"]), dev(c["buf"]["obs"]), dev(c["act"]),
                              dev(c["adv"]), dev(c["ret"]), c["l_mul"])
    l_q, g_q, _ = O.ppo_loss_grad(c["theta_pi"], c["theta_old"], c["buf"]["obs"], c["act"], c["adv"], c["ret"],
                                  c["l_mul"], lay, q=O.bf16_round)
    l_ref, g_ref, _ = O.ppo_loss_grad(c["theta_pi"], c["theta_old"], c["buf"]["obs"], c["act"], c["adv"], c["ret"],
                                      c["l_mul"], lay)
    got = g.cpu().numpy()[:lay.n_params]
    errs = per_tensor_rel_err(got, g_q, lay)
    print(name, "bf16 vs bf16-rounding oracle", {k: f"{v:.2e}" for k, v in errs.items()})
    print(name, "bf16 vs fp64 oracle (informative)", {k: f"{v:.2e}" for k, v in per_tensor_rel_err(got, g_ref, lay).items()})
    np.testing.assert_allclose(losses.cpu().numpy()[:4], l_q, rtol=1e-4, atol=1e-4)
    # 1024-row deep case: a single flipped ReLU/clip decision moves a gradient entry by ~1/sqrt(rows) = 3 %
    assert max(errs.values()) < (2e-3 if B >= 8192 else 2e-2), errs


# ------------------------------------------------------------------------------- epoch / minibatch scheduler
@pytest.mark.parametrize("n,seed,epoch", [(1, 0, 0), (2, 3, 1), (3, 3, 1), (1000, 1234, 0), (131072, 1234, 5),
                                          (100003, 2**63 + 5, 9)])
def test_minibatch_perm_bit_exact(n, seed, epoch):
    eng = _engine(3, (16,), 1, 128)
    got = eng.minibatch_perm(n, seed, epoch).cpu().numpy().astype(np.int64)
    ref = O.minibatch_perm(n, seed, epoch)
    assert np.array_equal(got, ref)
    assert np.array_equal(np.sort(got), np.arange(n))


@pytest.mark.parametrize("gemm_mode", [0, 1])
def test_update_epochs_whole_buffer_equals_update_steps_loop(gemm_mode):
    """n_epochs = UPDATE_STEPS, one whole-buffer minibatch, no shuffle  ==  Chief.py:64's loop of single updates."""
    T, N, obs, hidden, A = 256, 32, 376, (256, 256), 17
    lay = O.make_layout(obs, hidden, A)
    th_pi, th_old = O.make_params_pair(lay)
    buf = O.make_buffer(T, N, obs, A)
    eng = _engine(obs, hidden, A, T * N, gemm_mode=gemm_mode)
    d = {k: dev(v) for k, v in buf.items()}
    n = lay.n_params
    p1 = dev(th_pi); m1 = torch.zeros(n, device="cuda"); v1 = torch.zeros(n, device="cuda")
    l1 = []
    for t in (1, 2, 3, 4):
        _, g1 = eng.update(p1, dev(th_old), d["obs"], d["act"], d["rew"], d["predv"], d["done"], 0.99, 0.95, 0.7,
                           adam_m=m1, adam_v=v1, step=t, lr=2e-5 * 0.7)
        l1.append(g1[n:].clone())
    p2 = dev(th_pi); m2 = torch.zeros(n, device="cuda"); v2 = torch.zeros(n, device="cuda")
    l2, g2 = eng.update_epochs(p2, dev(th_old), d["obs"], d["act"], d["rew"], d["predv"], d["done"], 0.99, 0.95, 0.7,
                               m2, v2, step0=1, lr=2e-5 * 0.7, n_epochs=4)
    torch.cuda.synchronize()
    assert torch.equal(p1, p2) and torch.equal(m1, m2) and torch.equal(v1, v2)
    assert torch.equal(torch.stack(l1), l2) and torch.equal(g1, g2)
    assert not torch.equal(l2[0], l2[3])                 # the parameters did move between the epochs


@pytest.mark.parametrize("gemm_mode,mb,shuffle", [(0, 2048, True), (1, 2048, True), (1, 2048, False), (1, 3000, True)])
def test_update_epochs_minibatches_equal_manual_gather(gemm_mode, mb, shuffle):
    """The device scheduler (keyed permutation + row gather + per-minibatch loss_grad/Adam) against the same
    sequence driven from the host with torch indexing: bit-equal when every minibatch has the same row count."""
    T, N, obs, hidden, A = 128, 64, 376, (256, 256), 17
    B = T * N
    lay = O.make_layout(obs, hidden, A)
    th_pi, th_old = O.make_params_pair(lay)
    buf = O.make_buffer(T, N, obs, A)
    eng = _engine(obs, hidden, A, B, gemm_mode=gemm_mode)
    d = {k: dev(v) for k, v in buf.items()}
    n = lay.n_params
    adv, ret = eng.gae(d["rew"], d["predv"], d["done"], 0.99, 0.95)
    adv = adv.reshape(-1).clone(); ret = ret.reshape(-1).clone()
    p1 = dev(th_pi); m1 = torch.zeros(n, device="cuda"); v1 = torch.zeros(n, device="cuda")
    step = 1; l1 = []
    for e in range(2):
        order = torch.as_tensor(O.minibatch_perm(B, 77, e) if shuffle else np.arange(B), device="cuda")
        for r0 in range(0, B, mb):
            idx = order[r0:r0 + mb]
            _, g1 = eng.loss_grad(p1, dev(th_old), d["obs"][idx].contiguous(), d["act"][idx].contiguous(),
                                  adv[idx].contiguous(), ret[idx].contiguous(), 0.7)
            eng.adam_apply(p1, g1, m1, v1, step, 2e-5 * 0.7)
            l1.append(g1[n:].clone()); step += 1
    p2 = dev(th_pi); m2 = torch.zeros(n, device="cuda"); v2 = torch.zeros(n, device="cuda")
    l2, _ = eng.update_epochs(p2, dev(th_old), d["obs"], d["act"], d["rew"], d["predv"], d["done"], 0.99, 0.95, 0.7,
                              m2, v2, step0=1, lr=2e-5 * 0.7, n_epochs=2, minibatch_rows=mb, shuffle=shuffle, seed=77)
    torch.cuda.synchronize()
    l1 = torch.stack(l1)
    assert l2.shape == l1.shape == (2 * math.ceil(B / mb), 8)
    if B % mb == 0:
        assert torch.equal(l1, l2) and torch.equal(p1, p2) and torch.equal(m1, m2) and torch.equal(v1, v2)
    else:       # the short last minibatch uses the full minibatches' split-K plan: same numbers to fp32 rounding
        np.testing.assert_allclose(l2.cpu().numpy(), l1.cpu().numpy(), rtol=2e-5, atol=1e-6)
        assert rel_err(m2.cpu().numpy(), m1.cpu().numpy()) < 1e-4


def test_update_epochs_vs_oracle():
    """Shuffled minibatch epochs against the numpy restatement: per-minibatch losses within 1e-5 while the parameters
    move through 8 Adam steps (the fp32 update itself is compared where fp32 resolves it)."""
    T, N, obs, hidden, A = 64, 16, 17, (64, 64), 6
    B = T * N
    lay = O.make_layout(obs, hidden, A)
    th_pi, th_old = O.make_params_pair(lay)
    buf = O.make_buffer(T, N, obs, A)
    eng = _engine(obs, hidden, A, B)
    d = {k: dev(v) for k, v in buf.items()}
    n = lay.n_params
    lr = 1e-4
    p = dev(th_pi); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    l, _ = eng.update_epochs(p, dev(th_old), d["obs"], d["act"], d["rew"], d["predv"], d["done"], 0.99, 0.95, 0.7,
                             m, v, step0=1, lr=lr, n_epochs=2, minibatch_rows=256, shuffle=True, seed=5)
    th_ref, m_ref, v_ref, l_ref = O.update_epochs(th_pi, th_old, buf, 0.99, 0.95, 0.7, lay, lr, n_epochs=2,
                                                  minibatch_rows=256, shuffle=True, seed=5)
    lg = l.cpu().numpy()
    assert lg.shape == (8, 8)
    np.testing.assert_allclose(lg[:, :4], l_ref, rtol=3e-5, atol=3e-6)
    assert rel_err(m.cpu().numpy(), m_ref) < 1e-4
    upd = p.cpu().numpy().astype(np.float64) - th_pi
    assert rel_err(upd, th_ref - th_pi) < 2e-2            # |update| ~ 8e-4 on O(0.1) parameters: fp32 ulp / update


# ------------------------------------------------------------------------------- CategoricalPd path (distributions.py:124-159)
def _cat_engine(obs, hidden, ncat, max_rows, **kw):
    from dppo_b200 import _lib as L
    return _engine(obs, hidden, ncat, max_rows, pd_type=L.PD_CATEGORICAL, **kw)


def idev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.int32).cuda()


def test_categorical_methods_golden(golden_dir):
    """dppo_categorical ops against the reference class executed from source (tests/golden/categorical_ref.npz)."""
    g = np.load(os.path.join(golden_dir, "categorical_ref.npz"))
    eng = _cat_engine(3, (), 3, 64)
    p, q = dev(g["fix_p"][None]), dev(g["fix_q"][None])
    assert abs(float(eng.categorical("entropy", p)[0]) - float(g["fix_entropy"][0])) < 2e-6
    assert abs(float(eng.categorical("kl", p, q)[0]) - float(g["fix_kl"][0])) < 1e-6
    assert int(eng.categorical("mode", p)[0]) == int(g["fix_mode"][0])
    for k in range(3):
        assert abs(float(eng.categorical("neglogp", p, idev([k]))[0]) - float(g["fix_neglogp"][k][0])) < 2e-6
    assert float(eng.categorical("neglogp", p, idev([7]))[0]) == 0.0        # tf.one_hot of an out-of-range class is all zeros
    for n in (2, 6, 18):
        fp, fq, x = dev(g[f"n{n}_p"]), dev(g[f"n{n}_q"]), idev(g[f"n{n}_x"])
        np.testing.assert_allclose(eng.categorical("neglogp", fp, x).cpu().numpy(), g[f"n{n}_neglogp"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(eng.categorical("entropy", fp).cpu().numpy(), g[f"n{n}_entropy"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(eng.categorical("kl", fp, fq).cpu().numpy(), g[f"n{n}_kl"], rtol=1e-4, atol=2e-6)
        assert np.array_equal(eng.categorical("mode", fp).cpu().numpy(), g[f"n{n}_mode"])


def test_categorical_sampling_monte_carlo_identities():
    """The reference's self-check on its categorical fixture (distributions.py:260,269-295): Gumbel-max samples,
    entropy == -E[logp] and KL == -H - E_p[log q], 3 sigma; plus class frequencies == softmax."""
    pdparam = np.array([-.2, .3, .5]); Nn = 100000
    eng = _cat_engine(3, (), 3, Nn)
    M = dev(np.repeat(pdparam[None], Nn, 0))
    X = eng.categorical("sample", M, seed=42, offset=0)
    X2 = eng.categorical("sample", M, seed=42, offset=0); X3 = eng.categorical("sample", M, seed=42, offset=1)
    assert torch.equal(X, X2) and not torch.equal(X, X3)
    Xn = X.cpu().numpy(); Mn = np.repeat(pdparam[None], Nn, 0)
    ll = O.cat_logp(Mn, Xn)
    assert abs(O.cat_entropy(Mn).mean() + ll.mean()) < 3 * ll.std() / math.sqrt(Nn)
    q = pdparam + np.random.default_rng(0).standard_normal(3) * 0.1
    M2 = np.repeat(q[None], Nn, 0)
    ll2 = O.cat_logp(M2, Xn)
    kl_dev = float(eng.categorical("kl", M, dev(M2)).mean())
    assert abs(kl_dev - (-O.cat_entropy(Mn).mean() - ll2.mean())) < 3 * ll2.std() / math.sqrt(Nn) + 1e-6
    sm = np.exp(pdparam) / np.exp(pdparam).sum()
    freq = np.bincount(Xn, minlength=3) / Nn
    assert np.all(np.abs(freq - sm) < 4 * np.sqrt(sm * (1 - sm) / Nn))


def test_categorical_loss_matches_reference_ppo_lines_golden(golden_dir):
    """PPO.py:29-40 executed from the reference source over CategoricalPd heads and the reference's own int32 [B,1]
    action placeholder (PPO.py:26), fed through identity heads."""
    g = np.load(os.path.join(golden_dir, "categorical_ref.npz"))
    for tag in "abc":
        k = f"loss_{tag}_"
        flat, flat_old = g[k + "flat"], g[k + "flat_old"]
        B, n = flat.shape
        P = 1 + n
        obs = np.concatenate([g[k + "v"][:, None], flat, g[k + "v_old"][:, None], flat_old], 1).astype(np.float32)
        lay = O.make_layout(2 * P, (), n, pd="cat")
        th = []
        for base in (0, P):
            t = np.zeros(lay.n_params, np.float32); p = O.unflatten(t, lay)
            p["Wv"][base, 0] = 1.0
            for j in range(n):
                p["Wp"][base + 1 + j, j] = 1.0
            th.append(t)
        eng = _cat_engine(2 * P, (), n, B)
        losses, _ = eng.loss_grad(dev(th[0]), dev(th[1]), dev(obs), idev(g[k + "act"][:, None]), dev(g[k + "adv"]),
                                  dev(g[k + "etr"]), float(g[k + "l_mul"]))
        np.testing.assert_allclose(losses.cpu().numpy()[:4], g[k + "losses"], rtol=2e-5, atol=2e-6)


CAT_CASES = [("cartpole", 200, 4, 4, (16,), 2), ("atari-ram-ish", 128, 32, 128, (256, 256), 18),
             ("odd", 37, 5, 11, (24, 40), 5), ("no-hidden", 64, 4, 5, (), 3)]


@pytest.mark.parametrize("gemm_mode", [0, 1])
@pytest.mark.parametrize("name,T,N,obs,hidden,ncat", CAT_CASES)
def test_categorical_loss_grad_vs_oracle(name, T, N, obs, hidden, ncat, gemm_mode):
    c = O.make_case(T, N, obs, hidden, ncat, pd="cat", head_std=0.5, delta=0.1)
    lay = c["lay"]; B = T * N
    eng = _cat_engine(obs, hidden, ncat, B, gemm_mode=gemm_mode)
    losses, g = eng.loss_grad(dev(c["theta_pi"]), dev(c["theta_old"]), dev(c["buf"]["obs"]), idev(c["act"]),
                              dev(c["adv"]), dev(c["ret"]), c["l_mul"])
    l_ref, g_ref, aux = O.ppo_loss_grad(c["theta_pi"], c["theta_old"], c["buf"]["obs"], c["act"], c["adv"], c["ret"],
                                        c["l_mul"], lay)
    got_l = losses.cpu().numpy(); got_g = g.cpu().numpy()[:lay.n_params]
    np.testing.assert_allclose(got_l[:4], l_ref, rtol=TOL, atol=TOL)
    assert abs(got_l[4] - aux["mean_kl"]) < 1e-5 * max(1, aux["mean_kl"]) + 1e-7
    assert abs(got_l[5] - aux["clip_frac"]) < 1e-9 + 1.0 / B and got_l[7] == B
    errs = per_tensor_rel_err(got_g, g_ref, lay)
    print(name, "per-tensor rel err", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < (TOL if gemm_mode == 0 else 2e-5), errs
    assert aux["clip_frac"] > 0.01


def test_categorical_act_and_update_epochs():
    """dppo_act draws int32 classes (Gumbel max over the policy logits); the whole scheduler path runs on int32 actions
    (gather moves 4-byte words) and matches the host-driven sequence bit for bit."""
    T, N, obs, hidden, ncat = 64, 32, 8, (64, 64), 6
    B = T * N
    lay = O.make_layout(obs, hidden, ncat, pd="cat")
    th_pi, th_old = O.make_params_pair(lay, 1.0, 0.5, 0.05)
    buf = O.make_buffer(T, N, obs, ncat, pd="cat")
    eng = _cat_engine(obs, hidden, ncat, B)
    a, v = eng.act(dev(th_pi), dev(buf["obs"]), seed=3, offset=0)
    assert a.dtype == torch.int32 and a.shape == (B, 1) and int(a.min()) >= 0 and int(a.max()) < ncat
    v_ref, logits_ref, _ = O.mlp_forward(th_pi, buf["obs"], lay)
    assert rel_err(v.cpu().numpy(), v_ref) < TOL
    sm = np.exp(logits_ref - logits_ref.max(-1, keepdims=True)); sm /= sm.sum(-1, keepdims=True)
    freq = np.bincount(a.cpu().numpy()[:, 0], minlength=ncat) / B
    assert np.all(np.abs(freq - sm.mean(0)) < 5 * np.sqrt(0.25 / B))
    d = {k: (idev(x) if k == "act" else dev(x)) for k, x in buf.items()}
    n = lay.n_params
    adv, ret = eng.gae(d["rew"], d["predv"], d["done"], 0.99, 0.95)
    adv = adv.reshape(-1).clone(); ret = ret.reshape(-1).clone()
    p1 = dev(th_pi); m1 = torch.zeros(n, device="cuda"); v1 = torch.zeros(n, device="cuda")
    order = torch.as_tensor(O.minibatch_perm(B, 9, 0), device="cuda")
    for k, r0 in enumerate(range(0, B, 512)):
        idx = order[r0:r0 + 512]
        _, g1 = eng.loss_grad(p1, dev(th_old), d["obs"][idx].contiguous(), d["act"][idx].contiguous(),
                              adv[idx].contiguous(), ret[idx].contiguous(), 0.7)
        eng.adam_apply(p1, g1, m1, v1, k + 1, 1e-4)
    p2 = dev(th_pi); m2 = torch.zeros(n, device="cuda"); v2 = torch.zeros(n, device="cuda")
    eng.update_epochs(p2, dev(th_old), d["obs"], d["act"], d["rew"], d["predv"], d["done"], 0.99, 0.95, 0.7, m2, v2,
                      step0=1, lr=1e-4, n_epochs=1, minibatch_rows=512, shuffle=True, seed=9)
    torch.cuda.synchronize()
    assert torch.equal(p1, p2) and torch.equal(m1, m2)


# ------------------------------------------------------------------------------- the reference-named API on the real engine
@pytest.mark.parametrize("game,obs,pdw,discrete", [("Pendulum-v0", 3, 2, False), ("CartPole-v0", 4, 2, True)])
def test_reference_api_runs_on_engine(game, obs, pdw, discrete):
    """main.py / Chief / Worker / PPO (reference names) driving the CUDA engine end to end: 2 chief updates of
    UPDATE_STEPS=2, Box and Discrete action spaces; the worker-side loss fetch equals the oracle on the same feed."""
    import main as dppo_main
    # 200 steps: a Pendulum episode is 200 long and the reference pushes a batch only once an episode has ended (Worker.py:121,135)
    pd = dppo_main.build_parameter_dict(GAME=game, NUM_WORKERS=2, NUM_ENVS=4, MAX_EPOCH_STEPS=200, VERBOSE=False,
                                        UPDATE_STEPS=2, EPOCH_MAX=3, LOG_FILE_PATH="/tmp/dppo_test_logs")
    chief, workers = dppo_main.run(pd, max_updates=2)
    n = chief.ppo.pipara.flat.numel()
    assert n == obs * 16 + 16 + 16 + 1 + 16 * pdw + pdw
    assert chief.ppo._t == 4 and torch.isfinite(chief.ppo.pipara.flat).all()
    for wk in workers:
        assert torch.equal(wk.ppo.pipara.flat, chief.ppo.pipara.flat) and wk.ppo.discrete == discrete
    # one more loss evaluation through the placeholders, checked against the oracle
    w = workers[0].ppo
    rng = np.random.default_rng(0)
    B = 256
    s = rng.standard_normal((B, obs)).astype(np.float32)
    a = rng.integers(0, 2, (B, 1)).astype(np.int32) if discrete else rng.standard_normal((B, 1)).astype(np.float32)
    adv = rng.standard_normal((B, 1)).astype(np.float32); etr = rng.standard_normal((B, 1)).astype(np.float32)
    w.pipara.flat.add_(0.01 * torch.randn(n, device="cuda", generator=torch.Generator("cuda").manual_seed(1)))
    feed = {w.s: s, w.a: a, w.adv: adv, w.etr: etr, w.l_mul: 0.5}
    from RL.session import Session
    got = Session().run([w.policyLoss, w.valueLoss, w.entropyLoss, w.total_loss], feed)
    lay = O.make_layout(obs, (16,), 2 if discrete else 1, pd="cat" if discrete else "gauss")
    ref, _, _ = O.ppo_loss_grad(w.pipara.flat.cpu().numpy(), w.oldpipara.flat.cpu().numpy(), s, a, adv, etr, 0.5, lay)
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-6)


# ------------------------------------------------------------------------------- planes path: TMA-fed persistent K-major kernel
# ------------------------------------------------------------------------------- persistent fp16-split kernels (gemm_mode 3)
@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (256, 256, 64), (1000, 256, 376), (777, 35, 256), (130, 128, 100),
                                   (4096, 256, 256), (20000, 256, 376), (129, 200, 33), (300, 64, 17), (5000, 1024, 1024)])
def test_h_gemm_fp16_split(M, N, K):
    """tc_gemm_h.cu through dppo_dbg_gemm: persistent tcgen05 kind::f16 product with the 2-term fp16 split vs an fp64
    product, forward epilogue (bias + ReLU) and dX epilogue (fp32 mask)."""
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    eng = _engine(376, (256, 256), 17, max(M, 32768), gemm_mode=1)      # the weight image is taken from the workspace
    A = rng.standard_normal((M, K)).astype(np.float32)
    Bt = (rng.standard_normal((N, K)) / math.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    ref = A.astype(np.float64) @ Bt.astype(np.float64).T + bias
    got = eng.dbg_gemm(96, dev(A), dev(Bt), M, N, K, bias=dev(bias), relu=True).cpu().numpy()
    assert rel_err(got, np.maximum(ref, 0)) < 2e-6
    mask = rng.standard_normal((M, N)).astype(np.float32)
    ref2 = (A.astype(np.float64) @ Bt.astype(np.float64).T) * (mask > 0)
    got2 = eng.dbg_gemm(97, dev(A), dev(Bt), M, N, K, mask=dev(mask)).cpu().numpy()
    assert rel_err(got2, ref2) < 2e-6


@pytest.mark.parametrize("M,N,Kr", [(128, 256, 32), (376, 256, 1000), (256, 36, 2048), (120, 64, 333), (256, 256, 4096),
                                    (376, 256, 20000), (256, 128, 1500), (1024, 1024, 2048)])
def test_h_dw_fp16_split(M, N, Kr):
    """tc_dw_h.cu through dppo_dbg_gemm mode 99: C = X^T . G over the rows (split-K slices, fused column sums of G) vs fp64."""
    rng = np.random.default_rng(M + 3 * N + 7 * Kr)
    eng = _engine(376, (1024, 1024), 17, 4096, gemm_mode=1)
    X = rng.standard_normal((Kr, M)).astype(np.float32)
    G = rng.standard_normal((Kr, N)).astype(np.float32)       # the update feeds gradients pre-scaled to O(1) (capi.cu: grad_prescale)
    got, bias = eng.dbg_gemm(99, dev(X), dev(G), M, N, Kr, want_bias_out=True)
    ref = X.astype(np.float64).T @ G.astype(np.float64)
    assert rel_err(got.cpu().numpy(), ref) < 2e-6
    assert rel_err(bias.cpu().numpy(), G.astype(np.float64).sum(0)) < 2e-6


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (1000, 256, 376), (777, 256, 256), (129, 200, 40), (4096, 64, 256),
                                   (300, 128, 64), (20000, 256, 256), (3000, 1024, 1024)])
def test_h_gemm_fp16_planes(M, N, K):
    """tc_gemm_h.cu with the activation operand given as fp16 hi | lo planes (TMA writes the MMA operand tiles, no converter
    warps) and the result written as planes by TMA store -- the operand path of the update's hidden layers (modes 94 / 95)."""
    rng = np.random.default_rng(M * 5 + N * 3 + K)
    eng = _engine(376, (1024, 1024), 17, max(M, 8192), gemm_mode=1)
    A = rng.standard_normal((M, K)).astype(np.float32)
    Bt = (rng.standard_normal((N, K)) / math.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    ref = A.astype(np.float64) @ Bt.astype(np.float64).T + bias
    got = eng.dbg_gemm(94, dev(A), dev(Bt), M, N, K, bias=dev(bias), relu=True).cpu().numpy()
    assert rel_err(got, np.maximum(ref, 0)) < 2e-6
    mask = rng.standard_normal((M, N)).astype(np.float32)
    ref2 = (A.astype(np.float64) @ Bt.astype(np.float64).T) * (mask > 0)
    got2 = eng.dbg_gemm(95, dev(A), dev(Bt), M, N, K, mask=dev(mask)).cpu().numpy()
    assert rel_err(got2, ref2) < 2e-6
    # the planes hold exactly the split the converter warps apply to an fp32 operand: feeding the SAME values as fp32 gives the
    # same accumulators, so the two paths differ only by the final rounding of C to hi + lo / 2^11
    got_f32 = eng.dbg_gemm(96, dev(A), dev(Bt), M, N, K, bias=dev(bias), relu=True).cpu().numpy()
    assert rel_err(got, got_f32) < 5e-7


@pytest.mark.parametrize("mode", [101, 102, 103])
@pytest.mark.parametrize("M,N,Kr", [(128, 256, 32), (376, 256, 1000), (256, 40, 2048), (120, 64, 333), (256, 256, 4096),
                                    (376, 256, 20000), (256, 128, 1500), (1024, 1024, 2048)])
def test_h_dw_fp16_planes(M, N, Kr, mode):
    """tc_dw_h.cu with G (101), X (102) or both (103) given as fp16 hi | lo planes: TMA boxes are the MN-major operand atoms,
    the bias gradient is read back from the operand tiles."""
    rng = np.random.default_rng(M + 3 * N + 7 * Kr)
    eng = _engine(376, (1024, 1024), 17, 8192, gemm_mode=1)
    X = rng.standard_normal((Kr, M)).astype(np.float32)
    G = rng.standard_normal((Kr, N)).astype(np.float32)
    if Kr * max(M, N) > 8192 * 1024:
        pytest.skip("operand planes exceed the scratch buffers of this engine")
    got, bias = eng.dbg_gemm(mode, dev(X), dev(G), M, N, Kr, want_bias_out=True)
    ref = X.astype(np.float64).T @ G.astype(np.float64)
    assert rel_err(got.cpu().numpy(), ref) < 2e-6
    assert rel_err(bias.cpu().numpy(), G.astype(np.float64).sum(0)) < 2e-6
    # identical operand values as the fp32-fed kernel -> identical products
    got0, _ = eng.dbg_gemm(99, dev(X), dev(G), M, N, Kr, want_bias_out=True)
    assert torch.equal(got, got0)


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (1000, 256, 376), (777, 256, 256), (129, 200, 40), (4096, 64, 256),
                                   (300, 128, 64), (3000, 1024, 1024)])
def test_h_gemm_one_term_bf16(M, N, K):
    """The persistent kernels in their one-term mode (DPPO_GEMM_BF16 at config 4): bf16 operands, one MMA per product,
    double-buffered TMEM.  fp32 in / out (converters, modes 196 / 197): exact vs an fp64 product of the bf16-rounded operands;
    planes in / out (194 / 195): the result is additionally rounded to one bf16 plane."""
    rng = np.random.default_rng(M * 5 + N * 3 + K + 1)
    eng = _engine(376, (1024, 1024), 17, max(M, 8192), gemm_mode=1)
    A = rng.standard_normal((M, K)).astype(np.float32)
    Bt = (rng.standard_normal((N, K)) / math.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32); mask = rng.standard_normal((M, N)).astype(np.float32)
    ref = _bf16(A) @ _bf16(Bt).T
    got = eng.dbg_gemm(196, dev(A), dev(Bt), M, N, K, bias=dev(bias), relu=True).cpu().numpy()
    assert rel_err(got, np.maximum(ref + bias, 0)) < 3e-6
    got = eng.dbg_gemm(197, dev(A), dev(Bt), M, N, K, mask=dev(mask)).cpu().numpy()
    assert rel_err(got, ref * (mask > 0)) < 3e-6
    got = eng.dbg_gemm(194, dev(A), dev(Bt), M, N, K, bias=dev(bias), relu=True).cpu().numpy()
    want = np.maximum(ref + bias, 0)
    assert np.array_equal(got, _bf16(got).astype(np.float32)) and rel_err(got, want) < 4e-3        # one bf16 rounding of C
    got = eng.dbg_gemm(195, dev(A), dev(Bt), M, N, K, mask=dev(mask)).cpu().numpy()
    assert rel_err(got, ref * (mask > 0)) < 4e-3


@pytest.mark.parametrize("mode", [199, 201, 202, 203])
@pytest.mark.parametrize("M,N,Kr", [(128, 256, 32), (376, 256, 1000), (256, 40, 2048), (120, 64, 333), (256, 256, 4096),
                                    (1024, 1024, 2048)])
def test_h_dw_one_term_bf16(M, N, Kr, mode):
    """tc_dw_h.cu in one-term mode: bf16 operands (fp32 converted by the converter warps, or bf16 planes), fp32 accumulate."""
    rng = np.random.default_rng(M + 3 * N + 7 * Kr + 1)
    eng = _engine(376, (1024, 1024), 17, 8192, gemm_mode=1)
    X = rng.standard_normal((Kr, M)).astype(np.float32)
    G = rng.standard_normal((Kr, N)).astype(np.float32)
    got, bias = eng.dbg_gemm(mode, dev(X), dev(G), M, N, Kr, want_bias_out=True)
    assert rel_err(got.cpu().numpy(), _bf16(X).T @ _bf16(G)) < 3e-6
    g_seen = _bf16(G) if mode in (201, 203) else G.astype(np.float64)      # planes hold the rounded values
    assert rel_err(bias.cpu().numpy(), g_seen.sum(0)) < 3e-6


@pytest.mark.parametrize("obs,hidden,A,T,N", [(376, (256, 256), 17, 64, 48), (17, (256, 256), 6, 128, 16), (24, (64,), 4, 100, 9),
                                             (376, (1024, 1024, 1024, 1024), 17, 32, 16)])
def test_fp16x2_planes_equal_fp32_activations(obs, hidden, A, T, N, monkeypatch):
    """Hidden activations / gradients as fp16 hi | lo planes (default) against DPPO_H_PLANES=0 (fp32 matrices + converter
    warps): the operands of every product are the same numbers, so losses and weight gradients are bit-identical; bias gradients
    are summed from hi + lo / 2^11 instead of the fp32 value."""
    c = O.make_case(T, N, obs, hidden, A)
    lay = c["lay"]; B = T * N
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("DPPO_H_PLANES", flag)
        eng = _engine(obs, hidden, A, B, gemm_mode=3)
        losses, g = eng.loss_grad(dev(c["theta_pi"]), dev(c["theta_old"]), dev(c["buf"]["obs"]), dev(c["act"]),
                                  dev(c["adv"]), dev(c["ret"]), c["l_mul"])
        out[flag] = (losses.cpu().numpy().copy(), g.cpu().numpy()[:lay.n_params].copy())
    np.testing.assert_array_equal(out["1"][0], out["0"][0])
    for name, shape, off in lay.entries:
        n = int(np.prod(shape))
        a, b = out["1"][1][off:off + n], out["0"][1][off:off + n]
        if len(shape) == 2 and not (name == "W1" and obs % 4 != 0):      # (unaligned obs rows: layer 1's dW joins dh back to fp32)
            np.testing.assert_array_equal(a, b, err_msg=name)
        else:
            assert rel_err(a, b) < 2e-6, name


@pytest.mark.parametrize("name,T,N,obs,hidden,A", CASES)
def test_loss_grad_vs_oracle_fp16x2(name, T, N, obs, hidden, A):
    """The whole tower on the persistent fp16-split kernels (forward, dX; dW on the TMA-fed tf32 kernel), 1e-5 bar."""
    c = O.make_case(T, N, obs, hidden, A)
    lay = c["lay"]; B = T * N
    eng = _engine(obs, hidden, A, B, gemm_mode=3)
    losses, g = eng.loss_grad(dev(c["theta_pi"]), dev(c["theta_old"]), dev(c["buf"]["obs"]), dev(c["act"]),
                              dev(c["adv"]), dev(c["ret"]), c["l_mul"])
    l_ref, g_ref, aux = O.ppo_loss_grad(c["theta_pi"], c["theta_old"], c["buf"]["obs"], c["act"], c["adv"], c["ret"],
                                        c["l_mul"], lay)
    np.testing.assert_allclose(losses.cpu().numpy()[:4], l_ref, rtol=TOL, atol=TOL)
    errs = per_tensor_rel_err(g.cpu().numpy()[:lay.n_params], g_ref, lay)
    print(name, "fp16x2 per-tensor rel err", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < (3e-5 if max(hidden or (0,)) >= 1024 else TOL), errs


# ------------------------------------------------------------------------------- checkpoint / resume (tf_util.py:271-279)
def test_save_load_state_resume_equals_continuous(tmp_path):
    """4 updates in one go == 2 updates, dppo_save_state, fresh engine + dppo_load_state, 2 more updates: bit-identical
    parameters and Adam slots; a checkpoint of a different network is refused loudly."""
    T, N, obs, hidden, A = 64, 8, 17, (64, 64), 6
    c = O.make_case(T, N, obs, hidden, A)
    lay = c["lay"]; n = lay.n_params; buf = c["buf"]
    args = lambda: (dev(buf["obs"]), dev(c["act"]), dev(buf["rew"]), dev(buf["predv"]), dev(buf["done"]))   # noqa: E731

    def run(eng, p, old, m, v, steps):
        for t in steps:
            eng.update(p, old, *args(), 0.99, 0.95, 0.7, adam_m=m, adam_v=v, step=t, lr=1e-3)
        torch.cuda.synchronize()

    eng = _engine(obs, hidden, A, T * N)
    pa, old = dev(c["theta_pi"]), dev(c["theta_old"])
    ma = torch.zeros(n, device="cuda"); va = torch.zeros(n, device="cuda")
    run(eng, pa, old, ma, va, (1, 2, 3, 4))
    pb = dev(c["theta_pi"]); mb = torch.zeros(n, device="cuda"); vb = torch.zeros(n, device="cuda")
    run(eng, pb, old, mb, vb, (1, 2))
    path = str(tmp_path / "ckpt" / "state.dppo")
    os.makedirs(os.path.dirname(path))
    eng.save_state(path, pb, old, mb, vb, step=2)
    eng2 = _engine(obs, hidden, A, T * N)
    pc = torch.zeros(n, device="cuda"); oc = torch.zeros(n, device="cuda")
    mc = torch.zeros(n, device="cuda"); vc = torch.zeros(n, device="cuda")
    t = eng2.load_state(path, pc, oc, mc, vc)
    assert t == 2 and torch.equal(pc, pb) and torch.equal(oc, old) and torch.equal(mc, mb) and torch.equal(vc, vb)
    run(eng2, pc, oc, mc, vc, (t + 1, t + 2))
    assert torch.equal(pc, pa) and torch.equal(mc, ma) and torch.equal(vc, va)
    other = _engine(obs, (32,), A, 128)
    z = torch.zeros(other.n_params, device="cuda")
    with pytest.raises(RuntimeError, match="different network"):
        other.load_state(path, z, z.clone())
    with pytest.raises(RuntimeError, match="cannot open"):
        eng2.load_state(str(tmp_path / "missing.dppo"), pc, oc)


# ------------------------------------------------------------------------------- literal pairwise value loss (quirk D7)
def test_literal_value_loss_matches_reference_lines_golden(golden_dir):
    """PPO.py:29-40 executed from the reference source on the LITERAL rank-3 value head ([B,1,1] against etr [B,1],
    tests/golden/ppo_loss_ref.npz `*_losses_literal`): DPPO_VLOSS_LITERAL_PAIRWISE, one tower = the whole batch."""
    g = np.load(os.path.join(golden_dir, "ppo_loss_ref.npz"))
    for tag in "abc":
        flat, flat_old = g[f"{tag}_flat"], g[f"{tag}_flat_old"]
        B, A2 = flat.shape
        P = 1 + A2
        obs = np.concatenate([g[f"{tag}_v"][:, None], flat, g[f"{tag}_v_old"][:, None], flat_old], 1).astype(np.float32)
        lay, th_pi = _identity_head_params(P, False)
        _, th_old = _identity_head_params(P, True)
        eng = _engine(2 * P, (), A2 // 2, B, value_loss_mode=1)
        losses, _ = eng.loss_grad(dev(th_pi), dev(th_old), dev(obs), dev(g[f"{tag}_act"]), dev(g[f"{tag}_adv"]),
                                  dev(g[f"{tag}_etr"]), float(g[f"{tag}_l_mul"]))
        np.testing.assert_allclose(losses.cpu().numpy()[:4], g[f"{tag}_losses_literal"], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("gemm_mode", [0, 1])
def test_literal_value_loss_update_vs_oracle(gemm_mode):
    """The as-written mode through dppo_update: N towers of T rows (per-rollout [T,T] pairing, tower mean) and through
    dppo_loss_grad (one tower = all rows), losses and every gradient tensor against the oracle's closed form."""
    T, N, obs, hidden, A = 256, 8, 17, (64, 64), 6
    c = O.make_case(T, N, obs, hidden, A)
    lay = c["lay"]; n = lay.n_params; buf = c["buf"]; B = T * N
    eng = _engine(obs, hidden, A, B, gemm_mode=gemm_mode, value_loss_mode=1)
    adv, ret = O.gae_batch_vectorized(buf["rew"], buf["predv"], buf["done"], 0.99, 0.95)
    adv = adv.reshape(-1); ret = ret.reshape(-1)
    l_ref, g_ref, aux = O.ppo_loss_grad(c["theta_pi"], c["theta_old"], buf["obs"], c["act"], adv, ret, 0.7, lay,
                                        literal_groups=N)
    losses, g = eng.update(dev(c["theta_pi"]), dev(c["theta_old"]), dev(buf["obs"]), dev(c["act"]), dev(buf["rew"]),
                           dev(buf["predv"]), dev(buf["done"]), 0.99, 0.95, 0.7)
    got = losses.cpu().numpy()
    np.testing.assert_allclose(got[:4], l_ref, rtol=TOL, atol=TOL)
    assert abs(got[6] - aux["vclip_frac"]) < 1e-6
    # the pairwise value gradient is small against the policy terms: compare per tensor, relative to its max
    assert max(per_tensor_rel_err(g.cpu().numpy()[:n], g_ref, lay).values()) < 3 * TOL
    l1, g1, _ = O.ppo_loss_grad(c["theta_pi"], c["theta_old"], buf["obs"], c["act"], adv, ret, 0.7, lay, literal_groups=1)
    losses1, gg1 = eng.loss_grad(dev(c["theta_pi"]), dev(c["theta_old"]), dev(buf["obs"]), dev(c["act"]), dev(adv), dev(ret), 0.7)
    np.testing.assert_allclose(losses1.cpu().numpy()[:4], l1, rtol=TOL, atol=TOL)
    assert max(per_tensor_rel_err(gg1.cpu().numpy()[:n], g1, lay).values()) < 3 * TOL
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    with pytest.raises(RuntimeError, match="LITERAL_PAIRWISE"):
        eng.update_epochs(dev(c["theta_pi"]), dev(c["theta_old"]), dev(buf["obs"]), dev(c["act"]), dev(buf["rew"]),
                          dev(buf["predv"]), dev(buf["done"]), 0.99, 0.95, 0.7, m, v, step0=1, lr=1e-4, minibatch_rows=512)


# ------------------------------------------------------------------------------- Chief._train on the fused host-buffer call
def _chief_with_batches(update_steps, gemm_mode=0):
    """A Chief + 2 Workers on the HalfCheetah-shaped synthetic env with one queued rollout batch per worker."""
    import main as dppo_main
    from collections import deque
    from RL.Chief import Chief
    from RL.Worker import Worker
    from RL.session import Session, Coordinator
    import RL.PPO as rlppo
    rlppo._ENGINES.clear()
    pd = dppo_main.build_parameter_dict(GAME="HalfCheetah-v2", NUM_WORKERS=2, NUM_ENVS=4, MAX_EPOCH_STEPS=64, VERBOSE=False,
                                        UPDATE_STEPS=update_steps, HIDDEN=(64, 64), SEED=3, LOG_FILE_PATH="/tmp/dppo_test_logs")
    sess, coord, mem = Session(), Coordinator(), {}
    workers = []
    for i in range(2):
        mem["Worker_N%i" % i] = deque()
        workers.append(Worker("Worker_N%i" % i, pd, sess, mem, coord))
    chief = Chief("Chief", pd, sess, mem, coord, workers)
    rng = np.random.default_rng(11)
    rows = 64 * 4
    lay = O.make_layout(17, (64, 64), 6)
    th_pi, th_old = O.make_params_pair(lay)
    chief.ppo.pipara.flat.copy_(dev(th_pi))
    batches = []
    for i, wk in enumerate(workers):
        wk.ppo.pipara.flat.copy_(dev(th_pi)); wk.ppo.oldpipara.flat.copy_(dev(th_old))
        b = [rng.standard_normal((rows, 17)).astype(np.float32), rng.standard_normal((rows, 6)).astype(np.float32),
             rng.standard_normal((rows, 1)).astype(np.float32), rng.standard_normal((rows, 1)).astype(np.float32), 0.7,
             [1.0, -1.0, float(10 - i), 0.0, 0.0, 0.0, 0.0, 0.0, 1]]
        mem[wk.name].append(b)
        batches.append(b)
    return chief, workers, batches, lay, th_pi, th_old, pd


def test_chief_train_runs_the_fused_host_call_and_equals_the_engine_path():
    """Chief._train (Chief.py:54-64) -> ONE dppo_train_host call: parameters bit-identical to the same UPDATE_STEPS
    driven step by step through dppo_loss_grad + dppo_adam_apply on the concatenated towers, losses within 1e-5 of the oracle."""
    K = 3
    chief, workers, batches, lay, th_pi, th_old, pd = _chief_with_batches(K)
    eng = chief.ppo.backend
    eng.profile_reset(False)
    chief._train()
    torch.cuda.synchronize()
    assert chief.ppo._t == K and len(chief.ppo.step_losses) == K
    counts = {k: c for k, (c, _) in eng.profile_read().items()}
    # K steps on the device from one call (one rank: the Adam step runs inside finalize_kernel, no adam_kernel launch)
    assert counts["finalize"] == K and counts["loss"] == K and counts["adam"] == 0
    # the same K steps through the separate entry points
    s = dev(np.concatenate([b[0] for b in batches])); a = dev(np.concatenate([b[1] for b in batches]))
    adv = dev(np.concatenate([b[2] for b in batches]).reshape(-1)); etr = dev(np.concatenate([b[3] for b in batches]).reshape(-1))
    p = dev(th_pi); old = dev(th_old); n = lay.n_params
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    th = th_pi.astype(np.float64); m_ref = np.zeros(n); v_ref = np.zeros(n)
    for t in range(1, K + 1):
        l_ref, g_ref, _ = O.ppo_loss_grad(th.astype(np.float32), th_old, s.cpu().numpy(), a.cpu().numpy(), adv.cpu().numpy(),
                                          etr.cpu().numpy(), 0.7, lay)
        np.testing.assert_allclose(chief.ppo.step_losses[t - 1][:4], l_ref, rtol=2e-5, atol=2e-6)
        th, m_ref, v_ref = O.adam_tf(th, g_ref, m_ref, v_ref, t, pd["LEARNING_RATE"] * 0.7)
        _, g = eng.loss_grad(p, old, s, a, adv, etr, 0.7)
        eng.adam_apply(p, g, m, v, t, pd["LEARNING_RATE"] * 0.7)
    assert torch.equal(chief.ppo.pipara.flat, p) and torch.equal(chief.ppo._m, m) and torch.equal(chief.ppo._v, v)
