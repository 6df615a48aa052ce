#!/usr/bin/env python
"""BASELINE.json configs[4]: minibatch sweep 256..65536 on the Humanoid-shaped buffer at 1 or N GPUs
(python -m torch.distributed.run --nproc-per-node N ... tools/minibatch_sweep.py): per minibatch size the time of one
optimisation step (gather + fwd + loss + bwd + exchange + Adam, through dppo_update_epochs with the on-device
shuffle), env-steps/s, algorithmic HBM GB/s, and the gradient exchange alone (fused NVLink peer kernel and plain
ncclAllReduce) as latency and bus GB/s.  Rank 0 prints one JSON line per case."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tensorflow-dppo_b200"))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import torch
import torch.distributed as dist
from dppo_b200 import DppoEngine
from dppo_b200.synthetic import WORKLOADS, rollout_buffer, params_pair
from dppo_b200.dist import init_engine_comm, max_over_ranks

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
T, N, obs, act, hidden = WORKLOADS["humanoid-cfg3"]
B = T * N
eng = DppoEngine(obs, act, hidden, B, device=local, gemm_mode=1)
if world > 1:
    init_engine_comm(eng, rank, world)
buf = {k: v.to(dev) for k, v in rollout_buffer(T, N, obs, act, seed=1234 + rank).items()}
p_pi, p_old = (t.to(dev) for t in params_pair(eng.layout))
p0 = p_pi.clone()
m = torch.zeros_like(p_pi); v = torch.zeros_like(p_pi); g = eng.new_grad_buffer()


def timed(fn, k):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record(); torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / k
    return max_over_ranks(ms, device=dev) if world > 1 else ms


grad_bytes = (eng.n_params + 8) * 4
for mb in (256, 1024, 4096, 16384, 65536, B):
    n_mb = (B + mb - 1) // mb
    p_pi.copy_(p0); m.zero_(); v.zero_()
    st = {"t": 0}

    def epoch():
        eng.update_epochs(p_pi, p_old, buf["obs"], buf["act"], buf["rew"], buf["predv"], buf["done"], 0.99, 0.95, 0.7, m, v,
                          step0=st["t"] + 1, lr=1e-6, n_epochs=2, minibatch_rows=mb, shuffle=mb < B, seed=3, grad_out=g)
        st["t"] += 2 * n_mb
    reps = 2 if n_mb > 64 else 5
    eng.profile_reset(False)
    ms_epoch = timed(epoch, reps) / 2
    eng.profile_reset(True)
    epoch()
    prof = eng.profile_read(); eng.profile_reset(False)
    ex_ms = prof["allreduce"][1] / max(prof["allreduce"][0], 1) if world > 1 else 0.0
    ms_step = ms_epoch / n_mb
    if rank == 0:
        print(json.dumps({
            "case": f"humanoid-cfg3 minibatch {mb}", "n_gpus": world, "rows_per_gpu_per_step": mb, "steps_per_epoch": n_mb,
            "ms_per_step": ms_step, "env_steps_per_s": world * mb / ms_step * 1e3,
            "alg_HBM_GBps_per_gpu": mb * 1600 / ms_step / 1e6,
            "exchange_us": ex_ms * 1e3, "exchange_kind": "fused peer-memory allreduce + Adam" if world > 1 else None,
            "exchange_bus_GBps": (2 * (world - 1) / world * grad_bytes / (ex_ms * 1e-3) / 1e9) if ex_ms > 0 else None}), flush=True)

if world > 1:       # the plain NCCL allreduce of the same buffer, alone
    ms = timed(lambda: eng.allreduce_avg(g), 50)
    if rank == 0:
        print(json.dumps({"case": "ncclAllReduce(avg) of the flat gradient alone", "n_gpus": world, "bytes": grad_bytes,
                          "us": ms * 1e3, "bus_GBps": 2 * (world - 1) / world * grad_bytes / (ms * 1e-3) / 1e9}), flush=True)
    dist.destroy_process_group()
