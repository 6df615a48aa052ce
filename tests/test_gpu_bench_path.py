"""Oracle parity of the EXACT code path bench.py times: the tensor-core towers (gemm_mode 1 = tf32x3, 3 = fp16x2) at the
full BASELINE sizes (configs[2]: T=2048, N=64, obs 376, act 17, 2x256; configs[1]: T=2048, N=32, obs 17, act 6) through the
single-call entry points dppo_update (device buffers) and dppo_update_host (pinned host buffers), all T*N rows against
O.gae_batch_vectorized + O.ppo_loss_grad (fp64), followed by 3 chained Adam steps against O.adam_tf
(/root/reference/Worker.py:82-92, PPO.py:21-46,53; Chief.py:64).  Tolerance: 1e-5 of each tensor's max, BASELINE north_star.

The synthetic obs / act rows within fp32 rounding of a ReLU / clip / min / max switch are resampled first
(O.sanitize_for_update: the entry points compute adv / ret themselves, so only obs / act may be edited)."""
import numpy as np
import pytest
import torch

from oracle import dppo_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
GAMMA, LAM, L_MUL = 0.99, 0.95, 0.7
LR_CHAIN = 1e-8          # tiny: three Adam steps must not carry any row across a branch switch of the loss
MODES = {"tf32x3": 1, "fp16x2": 3}


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).cuda()


def rel_err(got, ref):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-30))


def per_tensor(got, ref, lay):
    out = {}
    for name, shape, off in lay.entries:
        n = int(np.prod(shape))
        out[name] = rel_err(got[off:off + n], ref[off:off + n])
    return out


def _oracle_chain(T, N, obs, hidden, A, steps=3):
    """Frozen buffer + the oracle's losses / gradients / Adam state of `steps` consecutive updates."""
    lay = O.make_layout(obs, hidden, A)
    th_pi, th_old = O.make_params_pair(lay)
    buf = O.make_buffer(T, N, obs, A)
    adv, ret = O.gae_batch_vectorized(buf["rew"], buf["predv"], buf["done"], GAMMA, LAM)
    adv = adv.reshape(-1); ret = ret.reshape(-1)
    buf["obs"], buf["act"], edited = O.sanitize_for_update(th_pi, th_old, buf["obs"], buf["act"], adv, ret, L_MUL, lay)
    th = th_pi.astype(np.float64); m = np.zeros(lay.n_params); v = np.zeros(lay.n_params)
    chain = []
    for t in range(1, steps + 1):
        losses, g, aux = O.ppo_loss_grad(th.astype(np.float32), th_old, buf["obs"], buf["act"], adv, ret, L_MUL, lay)
        th, m, v = O.adam_tf(th, g, m, v, t, LR_CHAIN)
        chain.append(dict(losses=losses, g=g, theta=th.copy(), m=m.copy(), v=v.copy(), clip_frac=aux["clip_frac"]))
    return dict(lay=lay, th_pi=th_pi, th_old=th_old, buf=buf, chain=chain, T=T, N=N, obs=obs, hidden=hidden, A=A,
                edited=edited, adv=adv)


@pytest.fixture(scope="module")
def cfg3():
    return _oracle_chain(2048, 64, 376, (256, 256), 17)


@pytest.fixture(scope="module")
def cfg2():
    return _oracle_chain(2048, 32, 17, (256, 256), 6)


def _run_chain(case, mode, entry):
    from dppo_b200 import DppoEngine
    lay = case["lay"]; T, N = case["T"], case["N"]; n = lay.n_params
    eng = DppoEngine(case["obs"], case["A"], case["hidden"], T * N, gemm_mode=MODES[mode])
    buf = case["buf"]
    p = dev(case["th_pi"]); old = dev(case["th_old"])
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    if entry == "update_host":
        pin = {k: torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for k, x in buf.items()}
    else:
        d = {k: dev(x) for k, x in buf.items()}
    assert case["chain"][0]["clip_frac"] > 0.01              # the clip branches are exercised at full size
    for t, ref in enumerate(case["chain"], start=1):
        if entry == "update_host":
            losses, g = eng.update_host(p, old, pin["obs"], pin["act"], pin["rew"], pin["predv"], pin["done"], GAMMA, LAM,
                                        L_MUL, adam_m=m, adam_v=v, step=t, lr=LR_CHAIN)
            losses = np.array(losses)
        else:
            l_t, g = eng.update(p, old, d["obs"], d["act"], d["rew"], d["predv"], d["done"], GAMMA, LAM, L_MUL,
                                adam_m=m, adam_v=v, step=t, lr=LR_CHAIN)
            losses = l_t.cpu().numpy()
        torch.cuda.synchronize()
        np.testing.assert_allclose(losses[:4], ref["losses"], rtol=TOL, atol=TOL, err_msg=f"{mode}/{entry} step {t}")
        assert losses[7] == T * N
        errs = per_tensor(g.cpu().numpy()[:n], ref["g"], lay)
        print(f"{mode}/{entry} step {t}: per-tensor rel err", {k: f"{e:.2e}" for k, e in errs.items()})
        assert max(errs.values()) < TOL, (mode, entry, t, errs)
        # Adam slots carry the chained gradients; parameters move by ~LR_CHAIN per step (resolved to one fp32 ulp)
        np.testing.assert_allclose(m.cpu().numpy(), ref["m"], rtol=0, atol=TOL * np.abs(ref["m"]).max())
        np.testing.assert_allclose(v.cpu().numpy(), ref["v"], rtol=0, atol=2 * TOL * np.abs(ref["v"]).max())
        np.testing.assert_allclose(p.cpu().numpy(), ref["theta"], rtol=1.2e-7, atol=1e-7)
    eng.close()


@pytest.mark.parametrize("entry", ["update", "update_host"])
@pytest.mark.parametrize("mode", ["tf32x3", "fp16x2"])
def test_full_cfg3_bench_path_vs_oracle(cfg3, mode, entry):
    """BASELINE configs[2] (the benchmarked workload), all 131,072 rows, the kernels bench.py times."""
    _run_chain(cfg3, mode, entry)


@pytest.mark.parametrize("entry", ["update", "update_host"])
@pytest.mark.parametrize("mode", ["tf32x3", "fp16x2"])
def test_full_cfg2_bench_path_vs_oracle(cfg2, mode, entry):
    """BASELINE configs[1]: HalfCheetah-shaped, all 65,536 rows."""
    _run_chain(cfg2, mode, entry)


@pytest.mark.parametrize("mode", ["simt", "tf32x3", "fp16x2"])
def test_update_host_ragged_rows_vs_oracle(mode):
    """T*N = 50,000 rows: not a multiple of the 128-row loss tile and larger than one H2D chunk, so the host path
    runs several chunks with a tapered, ragged tail (round-1 advisor finding: the tail must stay tile-aligned)."""
    from dppo_b200 import DppoEngine
    T, N, obs, hidden, A = 1000, 50, 17, (256, 256), 6
    lay = O.make_layout(obs, hidden, A)
    th_pi, th_old = O.make_params_pair(lay)
    buf = O.make_buffer(T, N, obs, A)
    adv, ret = O.gae_batch_vectorized(buf["rew"], buf["predv"], buf["done"], GAMMA, LAM)
    adv = adv.reshape(-1); ret = ret.reshape(-1)
    buf["obs"], buf["act"], _ = O.sanitize_for_update(th_pi, th_old, buf["obs"], buf["act"], adv, ret, L_MUL, lay)
    l_ref, g_ref, _ = O.ppo_loss_grad(th_pi, th_old, buf["obs"], buf["act"], adv, ret, L_MUL, lay)
    eng = DppoEngine(obs, A, hidden, T * N, gemm_mode={"simt": 0, **MODES}[mode])
    pin = {k: torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for k, x in buf.items()}
    losses, g = eng.update_host(dev(th_pi), dev(th_old), pin["obs"], pin["act"], pin["rew"], pin["predv"], pin["done"],
                                GAMMA, LAM, L_MUL)
    losses = np.array(losses)
    np.testing.assert_allclose(losses[:4], l_ref, rtol=TOL, atol=TOL)
    assert losses[7] == T * N
    assert max(per_tensor(g.cpu().numpy()[:lay.n_params], g_ref, lay).values()) < TOL
