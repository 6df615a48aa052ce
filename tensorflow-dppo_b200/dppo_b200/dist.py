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


def all_ranks_true(flag: bool, device=None, group=None) -> bool:
    """Collective AND of a per-rank decision.  Chief.check (Chief.py:19-28) trains only when every worker queued a
    batch; across ranks that decision must be unanimous, because _train contains the gradient exchange: a rank that
    skipped it while the others entered would hang the job (and the parameters would diverge)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return bool(flag)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


def stable_seed(*parts) -> int:
    """Process-independent 31-bit seed from strings / ints (python's hash() of a str is salted per process)."""
    import zlib
    h = 0
    for p in parts:
        h = zlib.crc32(str(p).encode("utf-8"), h)
    return h & 0x7FFFFFFF
