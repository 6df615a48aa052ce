#!/usr/bin/env python
"""Microbenchmark of the tensor-core GEMM kernels through dppo_dbg_gemm (CUDA events, L2-cold not enforced).
   python tools/tc_microbench.py            (on the GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tensorflow-dppo_b200"))
import torch
from dppo_b200 import DppoEngine

eng = DppoEngine(376, 17, (256, 256), 131072, gemm_mode=1)
shapes = [  # (mode, M, N, K, label)
    (48, 18944, 256, 4096, "img  one wave, long K"),
    (48, 131072, 256, 376, "img  fwd L1"),
    (48, 131072, 256, 256, "img  fwd L2"),
    (48, 131072, 35, 256, "img  fwd head"),
    (49, 131072, 256, 35, "img  dX head"),
    (49, 131072, 256, 256, "img  dX L2"),
    (16, 18944, 256, 4096, "reg  one wave, long K"),
    (16, 131072, 256, 256, "reg  fwd L2"),
    (18, 376, 256, 131072, "tn   dW1"),
    (18, 256, 256, 131072, "tn   dW2"),
    (18, 256, 35, 131072, "tn   dWh"),
    (0, 131072, 256, 256, "simt fwd L2"),
]
for mode, M, N, K, label in shapes:
    if mode == 18:
        A = torch.randn(K, M, device="cuda"); B = torch.randn(K, N, device="cuda")
    elif mode == 0:
        A = torch.randn(M, K, device="cuda"); B = torch.randn(K, N, device="cuda")
    else:
        A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    mask = torch.randn(M, N, device="cuda") if mode == 49 else None
    bias = torch.randn(N, device="cuda") if mode in (48, 16, 0) else None
    relu = mode in (48, 16, 0)
    for _ in range(3):
        eng.dbg_gemm(mode, A, B, M, N, K, bias=bias, mask=mask, relu=relu)
    torch.cuda.synchronize()
    reps = 10
    eng.profile_reset(True)                      # CUDA events around every library launch: GEMM kernels only, no packing
    for _ in range(reps):
        eng.dbg_gemm(mode, A, B, M, N, K, bias=bias, mask=mask, relu=relu)
    prof = eng.profile_read(); eng.profile_reset(False)
    fam = "tc_gemm" if mode >= 16 else {0: "gemm_nn", 1: "gemm_nt", 2: "gemm_tn"}[mode]
    us = prof[fam][1] * 1e3 / reps
    fl = 2.0 * M * N * K
    print(f"{label:26s} M={M:7d} N={N:4d} K={K:7d}  {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s (algorithmic)")
