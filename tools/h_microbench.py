#!/usr/bin/env python
"""Microbenchmark + time line of the persistent fp16-split kernel (tc_gemm_h.cu) next to the tf32 image kernel.
   python tools/h_microbench.py            (on the GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tensorflow-dppo_b200"))
import torch
from dppo_b200 import DppoEngine

eng = DppoEngine(376, 17, (256, 256), 131072, gemm_mode=1)
shapes = [  # (mode, M, N, K, label)
    (48, 131072, 256, 376, "tf32 fwd L1 (1 net)"),
    (96, 131072, 256, 376, "h    fwd L1 (1 net)"),
    (48, 131072, 256, 256, "tf32 fwd L2 (1 net)"),
    (96, 131072, 256, 256, "h    fwd L2 (1 net)"),
    (48, 131072, 35, 256, "tf32 fwd head"),
    (96, 131072, 35, 256, "h    fwd head"),
    (49, 131072, 256, 256, "tf32 dX L2"),
    (97, 131072, 256, 256, "h    dX L2 (fp32 mask)"),
    (49, 131072, 256, 36, "tf32 dX head"),
    (97, 131072, 256, 36, "h    dX head"),
    (48, 18944, 256, 4096, "tf32 one wave, long K"),
    (96, 18944, 256, 4096, "h    one wave, long K"),
    (96, 131072, 128, 256, "h    N=128"),
    (19, 376, 256, 131072, "tf32 dW L1 (TMA)"),
    (99, 376, 256, 131072, "h    dW L1"),
    (19, 256, 256, 131072, "tf32 dW L2 (TMA)"),
    (99, 256, 256, 131072, "h    dW L2"),
    (19, 256, 36, 131072, "tf32 dW head (TMA)"),
    (99, 256, 36, 131072, "h    dW head"),
]
for mode, M, N, K, label in shapes:
    if mode in (19, 99):
        A = torch.randn(K, M, device="cuda"); B = torch.randn(K, N, device="cuda")
    else:
        A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    mask = torch.randn(M, N, device="cuda") if mode in (49, 97) else None
    bias = torch.randn(N, device="cuda") if mode in (48, 96) else None
    relu = mode in (48, 96)
    for _ in range(3):
        eng.dbg_gemm(mode, A, B, M, N, K, bias=bias, mask=mask, relu=relu)
    torch.cuda.synchronize()
    reps = 10
    eng.profile_reset(True)
    for _ in range(reps):
        eng.dbg_gemm(mode, A, B, M, N, K, bias=bias, mask=mask, relu=relu)
    prof = eng.profile_read(); eng.profile_reset(False)
    us = prof["tc_gemm"][1] * 1e3 / reps
    fl = 2.0 * M * N * K
    print(f"{label:26s} M={M:7d} N={N:4d} K={K:7d}  {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s (algorithmic)", flush=True)

# time line of CTA 0 (clock64): per tile [mma start, mma issued, epilogue sees accumulator, TMEM released, written out]
for (M, N, K) in ((131072, 256, 376), (131072, 256, 256)):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    tl = torch.zeros(128, dtype=torch.int64, device="cuda")
    eng.dbg_gemm(98, A, B, M, N, K, bias=torch.zeros(N, device="cuda"), relu=True, aux_out=tl)
    torch.cuda.synchronize()
    t = tl.cpu().numpy().reshape(16, 8)
    t0 = t[0, 0]
    print(f"time line M={M} N={N} K={K} (cycles since the first tile's MMA start; tiles of CTA 0)")
    for i in range(7):
        if t[i, 0] == 0:
            break
        print(f"  tile {i}: mma_start {t[i,0]-t0:7d}  mma_issued {t[i,1]-t0:7d}  acc_seen {t[i,2]-t0:7d}  tmem_free {t[i,3]-t0:7d}  "
              f"written {t[i,4]-t0:7d}   | tile period {t[i+1,0]-t[i,0] if t[i+1,0] else 0}")
