"""BASELINE configs[0] plumbing: Pendulum-shaped, 1 chief + 2 workers, through the reference-named
API surface (RL.Chief / RL.Worker / RL.PPO / main.run) WITHOUT a GPU.  The product has no CPU
backend; this test injects a test double that answers the engine calls with the oracle."""
import numpy as np
import pytest
import torch

from oracle import dppo_oracle as O


class OracleBackend:
    """Test double with DppoEngine's call surface, computing on the CPU oracle."""
    device = "cpu"
    world = 1

    def __init__(self, obs, act, hidden, pd="gauss"):
        self.lay = O.make_layout(obs, hidden, act, pd)
        self.hp = O.HParams()
        self.calls = []

    def gae(self, rew, predv, done, gamma, lam, **kw):
        self.calls.append("gae")
        a, r = O.gae_batch(rew.numpy(), predv.numpy(), done.numpy(), gamma, lam)
        return torch.from_numpy(a), torch.from_numpy(r)

    def act(self, params, obs, seed, offset):
        self.calls.append("act")
        v, flat, _ = O.mlp_forward(params.numpy(), obs.numpy(), self.lay)
        if self.lay.pd == "cat":
            u = np.random.default_rng(seed + offset).random(flat.shape)
            return (torch.from_numpy(O.cat_sample(flat, u).astype(np.int32)[:, None]),
                    torch.from_numpy(v.astype(np.float32)))
        eps = np.random.default_rng(seed + offset).standard_normal((obs.shape[0], self.lay.act_dim))
        return torch.from_numpy(O.dg_sample(flat, eps).astype(np.float32)), torch.from_numpy(v.astype(np.float32))

    def loss_grad(self, p_pi, p_old, s, a, adv, etr, l_mul):
        self.calls.append("loss_grad")
        losses, g, aux = O.ppo_loss_grad(p_pi.numpy(), p_old.numpy(), s.numpy(), a.numpy(), adv.numpy(), etr.numpy(),
                                         l_mul, self.lay, self.hp)
        tail = np.concatenate([losses, [aux["mean_kl"], aux["clip_frac"], aux["vclip_frac"], s.shape[0]]])
        buf = torch.from_numpy(np.concatenate([g, tail]).astype(np.float32))
        return buf[self.lay.n_params:], buf

    def sync_old(self, p_pi, p_old):
        self.calls.append("sync_old"); p_old.copy_(p_pi)

    def adam_apply(self, params, grad, m, v, step, lr, **kw):
        self.calls.append("adam")
        n = self.lay.n_params
        th, m2, v2 = O.adam_tf(params.numpy(), grad.numpy()[:n], m.numpy(), v.numpy(), step, lr)
        params.copy_(torch.from_numpy(th.astype(np.float32)))
        m.copy_(torch.from_numpy(m2.astype(np.float32))); v.copy_(torch.from_numpy(v2.astype(np.float32)))

    def allreduce_avg(self, buf):
        return buf


def test_reference_api_surface_and_two_updates():
    import main as dppo_main
    from RL.PPO import PPO
    from RL.Model import Model
    from RL.Others.distributions import make_pdtype, DiagGaussianPdType, DiagGaussianPd, Pd, PdType
    be = OracleBackend(3, 1, (16,))
    pd = dppo_main.build_parameter_dict(BACKEND=be, NUM_WORKERS=2, NUM_ENVS=2, MAX_EPOCH_STEPS=200, VERBOSE=False,
                                        UPDATE_STEPS=2, EPOCH_MAX=3)
    # the reference's keys are all present (main.py:12-29)
    for k in ("GAME", "LEARNING_RATE", "ENTCOEFF", "VCOEFF", "CLIP_PARAM", "GAMMA", "LAM", "SCHEDULE",
              "MAX_AC_EXP_RATE", "MIN_AC_EXP_RATE", "AC_EXP_PERCENTAGE", "UPDATE_STEPS", "MAX_EPOCH_STEPS",
              "EPOCH_MAX", "NUM_WORKERS", "LOG_FILE_PATH"):
        assert k in pd
    chief, workers = dppo_main.run(pd, max_updates=2)
    w = workers[0].ppo
    for attr in ("s", "l_mul", "pipredv", "pipara", "oldpipara", "ca", "a", "adv", "etr", "policyLoss",
                 "entropyLoss", "valueLoss", "total_loss", "summary_op", "gradient", "sync_pis"):
        assert hasattr(w, attr), attr
    assert w.action is w.a and w.act is w.a and w.advantage is w.adv and w.estimatedReturn is w.etr and w.str is w.etr
    assert hasattr(chief.ppo, "train") and hasattr(chief.ppo, "_average_gradients")
    for m in ("check", "_train", "_update_local_workers_weight", "_get_data_size", "_logs_writer", "act"):
        assert callable(getattr(chief, m))
    assert callable(workers[0].work) and callable(workers[0].act)
    # parameter list order and shapes = Model.py:17 (W1 [3,16], b1, Wv, bv, Wp, bp): 115 parameters
    assert [tuple(p.shape) for p in chief.ppo.pipara] == [(3, 16), (16,), (16, 1), (1,), (16, 2), (2,)]
    assert chief.ppo.pipara.flat.numel() == 115
    # two updates of UPDATE_STEPS=2 Adam steps happened, workers hold the chief's weights again
    assert be.calls.count("adam") == 4 and chief.ppo._t == 4
    for wk in workers:
        assert torch.equal(wk.ppo.pipara.flat, chief.ppo.pipara.flat)
    assert np.isfinite(chief.ppo.last_losses.numpy()).all()
    a = chief.act(np.zeros(3, np.float32))
    assert np.shape(a) == (1,)
    # distribution surface
    pt = make_pdtype(workers[0].env.action_space)
    assert isinstance(pt, DiagGaussianPdType) and isinstance(pt, PdType) and pt.param_shape() == [2] and pt.sample_shape() == [1]
    assert pt.pdclass() is DiagGaussianPd and issubclass(DiagGaussianPd, Pd)
    with pytest.raises(RuntimeError, match="no CPU path"):
        DiagGaussianPd(torch.zeros(4, 2))
    # Model.FC keeps the reference signature incl. the `state=` spelling of PPO.py:21-22
    _, _, para = Model().FC("x", state=w.s, action_space=workers[0].env.action_space, device="cpu")
    W1 = list(para)[0].numpy()
    np.testing.assert_allclose(np.sqrt((W1 ** 2).sum(0)), 0.01, rtol=1e-5)       # normc(0.01), Model.py:10


def test_cartpole_default_game_discrete_actions():
    """The reference's committed default (main.py:13 'CartPole-v0'): Discrete(2) -> CategoricalPd (distributions.py:236-237),
    int32 [B,1] action placeholder (PPO.py:26), epsilon-greedy exploration in Worker.act (Worker.py:141-152)."""
    import main as dppo_main
    from RL.Others.distributions import make_pdtype, CategoricalPdType, CategoricalPd
    be = OracleBackend(4, 2, (16,), pd="cat")
    pd = dppo_main.build_parameter_dict(GAME='CartPole-v0', BACKEND=be, NUM_WORKERS=2, NUM_ENVS=2, MAX_EPOCH_STEPS=100,
                                        VERBOSE=False, UPDATE_STEPS=2, EPOCH_MAX=3)
    chief, workers = dppo_main.run(pd, max_updates=2)
    w = workers[0]
    assert w.ppo.discrete and w.ppo.act_dim == 2 and w.ppo.act_cols == 1 and w.ppo.a.shape == [None, 1]
    # Model.py:14: the policy head is param_shape()[0] = ncat wide -> [W1 [4,16], b1, Wv, bv, Wp [16,2], bp [2]]
    assert [tuple(p.shape) for p in chief.ppo.pipara] == [(4, 16), (16,), (16, 1), (1,), (16, 2), (2,)]
    assert be.calls.count("adam") == 4
    batch = None
    for wk in workers:
        assert torch.equal(wk.ppo.pipara.flat, chief.ppo.pipara.flat)
    a, v = w.act(w.env.reset())
    assert a.dtype == np.int32 and a.shape == (2, 1) and set(np.unique(a)) <= {0, 1}
    pt = make_pdtype(w.env.action_space)
    assert isinstance(pt, CategoricalPdType) and pt.param_shape() == [2] and pt.sample_shape() == []
    assert pt.pdclass() is CategoricalPd
    with pytest.raises(RuntimeError, match="no CPU path"):
        CategoricalPd(torch.zeros(4, 2))
    assert np.isfinite(chief.ppo.last_losses.numpy()).all()


def test_filewriter_writes_tensorboard_event_file(tmp_path):
    """RL.session.FileWriter (tf.summary.FileWriter, main.py:46) emits a real event file: TensorBoard's own loader
    reads back the scalars of PPO.summary_op (PPO.py:41-45)."""
    from RL.session import FileWriter
    w = FileWriter(str(tmp_path))
    rows = [(1, dict(entropyLoss=-0.0144, policyLoss=0.0061, valueLoss=3.937, total_loss=3.929)),
            (2, dict(entropyLoss=-0.0143, policyLoss=0.0055, valueLoss=3.5, total_loss=3.49))]
    for step, sc in rows:
        w.add_summary(sc, step)
    w.close()
    loader_mod = pytest.importorskip("tensorboard.backend.event_processing.event_file_loader")
    got = []
    for ev in loader_mod.EventFileLoader(w.event_path).Load():
        if ev.HasField("summary") and len(ev.summary.value):
            vals = {}
            for v in ev.summary.value:
                # newer TensorBoard loaders migrate simple_value into a rank-0 float tensor
                vals[v.tag] = v.simple_value if v.HasField("simple_value") else float(v.tensor.float_val[0])
            got.append((ev.step, vals))
    assert [s for s, _ in got] == [1, 2]
    for (_, ref), (_, vals) in zip(rows, got):
        assert set(vals) == set(ref)
        for k in ref:
            assert abs(vals[k] - ref[k]) < 1e-6
