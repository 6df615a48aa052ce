"""Host-side multi-GPU plumbing: rank sharding of rollouts and communicator bootstrap.
One process per GPU; torch.distributed (nccl on GPUs, gloo in CPU tests) only carries the
128-byte NCCL unique id and the timing reductions -- the gradient mean itself is
dppo_allreduce_avg inside the library."""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_rollouts(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the rollout columns rank owns under STRONG scaling (n_total fixed).
    Rollouts are independent (per-rollout GAE and normalisation, Worker.py:82-92), so the only
    cross-rank coupling is the parameter-shaped gradient mean."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} of {world}")
    base, rem = divmod(n_total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def weighted_mean_coefficients(counts):
    """Weights that turn ncclAvg over ranks into the mean over all rows when ranks hold unequal row
    counts: scale rank r's gradient by counts[r]*world/sum(counts) before the allreduce."""
    total = float(sum(counts)); world = len(counts)
    return [c * world / total for c in counts]


def exchange_unique_id(make_id, rank: int, world: int, group=None) -> bytes:
    """rank 0 creates the NCCL unique id (make_id()), everyone receives the same 128 bytes."""
    payload = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0, group=group)
    assert isinstance(payload[0], (bytes, bytearray)) and len(payload[0]) == 128
    return bytes(payload[0])


def init_engine_comm(engine, rank: int, world: int, group=None):
    uid = exchange_unique_id(type(engine).unique_id, rank, world, group)
    engine.comm_init(rank, world, uid)
    return uid


def max_over_ranks(value: float, device=None, group=None) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
