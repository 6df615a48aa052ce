"""Host-side multi-rank logic on CPU (gloo, world_size 2): rollout sharding, unique-id exchange and
the identity  allreduce-average of per-rank tower gradients == PPO._average_gradients (PPO.py:55-65)
== gradient of the concatenated batch (equal-sized towers)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dppo_oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dppo_b200.dist import shard_rollouts, exchange_unique_id, max_over_ranks
    uid = exchange_unique_id(lambda: bytes(range(128)), rank, world)
    assert uid == bytes(range(128))
    # Chief.check's "train now?" must be unanimous (the train step contains the exchange): collective AND
    from dppo_b200.dist import all_ranks_true
    assert all_ranks_true(True) is True
    assert all_ranks_true(rank == 0) is False          # one rank without data -> nobody trains
    assert all_ranks_true(False) is False
    assert max_over_ranks(float(rank + 1)) == float(world)
    T, N, obs, hidden, A = 64, 8, 11, (16, 12), 3
    c = O.make_case(T, N, obs, hidden, A)
    lo, hi = shard_rollouts(N, rank, world)
    cols = np.arange(lo, hi)
    rows = (np.arange(T)[:, None] * N + cols[None, :]).reshape(-1)           # row = t*N + n
    losses, g, _ = O.ppo_loss_grad(c["theta_pi"], c["theta_old"], c["buf"]["obs"][rows], c["act"][rows],
                                   c["adv"][rows], c["ret"][rows], c["l_mul"], c["lay"])
    buf = torch.from_numpy(np.concatenate([g, losses]))
    dist.all_reduce(buf); buf /= world                                        # what ncclAvg does on the GPU path
    np.save(os.path.join(out_dir, f"avg_{rank}.npy"), buf.numpy())
    np.save(os.path.join(out_dir, f"own_{rank}.npy"), np.concatenate([g, losses]))
    dist.destroy_process_group()


def test_two_rank_gradient_mean_equals_tower_mean(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    avg = [np.load(tmp_path / f"avg_{r}.npy") for r in range(world)]
    own = [np.load(tmp_path / f"own_{r}.npy") for r in range(world)]
    np.testing.assert_array_equal(avg[0], avg[1])                             # every rank applies the same update
    np.testing.assert_allclose(avg[0], O.tower_mean(own), rtol=1e-14)
    T, N, obs, hidden, A = 64, 8, 11, (16, 12), 3
    c = O.make_case(T, N, obs, hidden, A)
    losses, g, _ = O.ppo_loss_grad(c["theta_pi"], c["theta_old"], c["buf"]["obs"], c["act"], c["adv"], c["ret"],
                                   c["l_mul"], c["lay"])
    np.testing.assert_allclose(avg[0][:-4], g, rtol=1e-9, atol=1e-14)         # == whole-batch gradient
    np.testing.assert_allclose(avg[0][-4:], losses, rtol=1e-12)


def test_shard_rollouts_partitions():
    from dppo_b200.dist import shard_rollouts, weighted_mean_coefficients
    for n in (1, 7, 64, 65, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_rollouts(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            w = weighted_mean_coefficients(sizes)
            assert abs(sum(wi * 1.0 for wi in w) / world - 1.0) < 1e-12 or n < world
    with pytest.raises(ValueError):
        shard_rollouts(8, 2, 2)


def test_stable_seed_is_process_independent_and_rank_dependent():
    """Seeds come from a stable hash (str hash() is salted per process) and differ per rank (advisor finding, round 1)."""
    import subprocess, sys
    from dppo_b200.dist import stable_seed
    code = ("import sys; sys.path.insert(0, %r); from dppo_b200.dist import stable_seed; "
            "print(stable_seed('Worker_N0', 0, 0), stable_seed('Worker_N0', 0, 1))" % os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tensorflow-dppo_b200"))
    outs = {subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PYTHONHASHSEED=str(h))).stdout.strip()
            for h in (1, 2)}
    assert len(outs) == 1
    a, b = (int(x) for x in outs.pop().split())
    assert a == stable_seed("Worker_N0", 0, 0) and a != b and 0 <= a < 2 ** 31
