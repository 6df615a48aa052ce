"""The fused NVLink peer-memory exchange + Adam kernel (csrc/comm.cu: peer_allreduce_adam_kernel), the kernel every
multi-GPU number runs on, proved on ONE device: two processes share cuda:0, bootstrap the exchange through
dppo_comm_peer_export / dppo_comm_peer_import (IPC handle blobs moved through files, no NCCL -- NCCL refuses two ranks
on one GPU), and run three consecutive dppo_update calls each on its own shard of rollouts.  Checked against the oracle's
tower mean (/root/reference/PPO.py:55-65) followed by TF-Adam (PPO.py:53): per-step averaged gradient and losses, Adam
slots, parameters, and bit-identical results on both ranks."""
import os
import time

import numpy as np
import pytest
import torch

from oracle import dppo_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
T, N, OBS, HIDDEN, A = 128, 16, 17, (64, 64), 6
GAMMA, LAM, L_MUL, LR = 0.99, 0.95, 0.7, 1e-8
STEPS = 3


def _shard(rank, world):
    from dppo_b200.dist import shard_rollouts
    lo, hi = shard_rollouts(N, rank, world)
    rows = (np.arange(T)[:, None] * N + np.arange(lo, hi)[None, :]).reshape(-1)
    return lo, hi, rows


def _case():
    lay = O.make_layout(OBS, HIDDEN, A)
    th_pi, th_old = O.make_params_pair(lay)
    buf = O.make_buffer(T, N, OBS, A)
    adv, ret = O.gae_batch_vectorized(buf["rew"], buf["predv"], buf["done"], GAMMA, LAM)
    buf["obs"], buf["act"], _ = O.sanitize_for_update(th_pi, th_old, buf["obs"], buf["act"], adv.reshape(-1),
                                                      ret.reshape(-1), L_MUL, lay)
    return lay, th_pi, th_old, buf, adv.reshape(-1), ret.reshape(-1)


def _worker(rank, world, out_dir, gemm_mode):
    torch.cuda.set_device(0)                                    # BOTH ranks on the same device
    from dppo_b200 import DppoEngine
    lay, th_pi, th_old, buf, _, _ = _case()
    lo, hi, rows = _shard(rank, world)
    eng = DppoEngine(OBS, A, HIDDEN, len(rows), device=0, gemm_mode=gemm_mode)
    blob = eng.comm_peer_export(rank, world)
    tmp = os.path.join(out_dir, f"blob_{rank}.tmp")
    with open(tmp, "wb") as f:
        f.write(blob)
    os.rename(tmp, os.path.join(out_dir, f"blob_{rank}.bin"))
    blobs = []
    for r in range(world):
        path = os.path.join(out_dir, f"blob_{r}.bin")
        t0 = time.time()
        while not os.path.exists(path):
            assert time.time() - t0 < 120, "peer handle never arrived"
            time.sleep(0.01)
        blobs.append(open(path, "rb").read())
    eng.comm_peer_import(blobs)
    assert eng.world == world
    d = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).cuda()   # noqa: E731
    p = d(th_pi); old = d(th_old); n = lay.n_params
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    rb = {k: d(buf[k][:, lo:hi]) for k in ("rew", "predv", "done")}
    obs_r, act_r = d(buf["obs"][rows]), d(buf["act"][rows])
    for step in range(1, STEPS + 1):
        _, g = eng.update(p, old, obs_r, act_r, rb["rew"], rb["predv"], rb["done"], GAMMA, LAM, L_MUL, adam_m=m, adam_v=v,
                          step=step, lr=LR)
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, f"g_{rank}_{step}.npy"), g.cpu().numpy())
    np.save(os.path.join(out_dir, f"p_{rank}.npy"), p.cpu().numpy())
    np.save(os.path.join(out_dir, f"m_{rank}.npy"), m.cpu().numpy())
    np.save(os.path.join(out_dir, f"v_{rank}.npy"), v.cpu().numpy())
    eng.close()


@pytest.mark.parametrize("gemm_mode", [0, 1])
def test_peer_exchange_adam_two_ranks_one_gpu(tmp_path, gemm_mode):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, str(tmp_path), gemm_mode), nprocs=world, join=True)
    lay, th_pi, th_old, buf, adv, ret = _case()
    n = lay.n_params
    th = th_pi.astype(np.float64); m = np.zeros(n); v = np.zeros(n)
    for step in range(1, STEPS + 1):
        towers = []
        for r in range(world):
            _, _, rows = _shard(r, world)
            losses, g, _ = O.ppo_loss_grad(th.astype(np.float32), th_old, buf["obs"][rows], buf["act"][rows], adv[rows],
                                           ret[rows], L_MUL, lay)
            towers.append(np.concatenate([g, losses]))
        ref = O.tower_mean(towers)                                              # PPO.py:55-65
        th, m, v = O.adam_tf(th, ref[:n], m, v, step, LR)                       # PPO.py:53
        g0 = np.load(tmp_path / f"g_0_{step}.npy"); g1 = np.load(tmp_path / f"g_1_{step}.npy")
        np.testing.assert_array_equal(g0, g1)                                   # rank-order sums: bit-identical on every rank
        for name, shape, off in lay.entries:
            k = int(np.prod(shape))
            err = np.max(np.abs(g0[off:off + k] - ref[off:off + k])) / np.max(np.abs(ref[off:off + k]))
            assert err < TOL, (step, name, err)
        np.testing.assert_allclose(g0[n:n + 4], ref[n:], rtol=TOL, atol=TOL)    # the loss scalars ride in the same exchange
    for what, ref_x, rt, at in (("p", th, 1.2e-7, 1e-7), ("m", m, 0, TOL * np.abs(m).max()), ("v", v, 0, 2 * TOL * np.abs(v).max())):
        x0 = np.load(tmp_path / f"{what}_0.npy"); x1 = np.load(tmp_path / f"{what}_1.npy")
        np.testing.assert_array_equal(x0, x1)
        np.testing.assert_allclose(x0, ref_x, rtol=rt, atol=at)
