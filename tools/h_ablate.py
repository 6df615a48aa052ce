#!/usr/bin/env python
"""One shape of the persistent fp16-split kernel (mode 96/97) timed + its CTA-0 time line; DPPO_H_DBG selects an ablation.
   python tools/h_ablate.py [M N K]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tensorflow-dppo_b200"))
import torch
from dppo_b200 import DppoEngine
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (131072, 256, 376)
eng = DppoEngine(376, 17, (256, 256), 131072, gemm_mode=1)
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); bias = torch.randn(N, device="cuda")
PL = os.environ.get("H_PLANES", "0") == "1"          # activation operand and result as fp16 hi | lo planes (modes 94 / 92)
m_run, m_tl = (94, 92) if PL else (96, 98)
if os.environ.get("H_TERMS", "2") == "1":            # one-term (bf16) mode of the same kernels
    m_run, m_tl = m_run + 100, m_tl + 100
for _ in range(3):
    eng.dbg_gemm(m_run, A, B, M, N, K, bias=bias, relu=True)
torch.cuda.synchronize()
eng.profile_reset(True)
for _ in range(10):
    eng.dbg_gemm(m_run, A, B, M, N, K, bias=bias, relu=True)
us = eng.profile_read()["tc_gemm"][1] * 100
tl = torch.zeros(128, dtype=torch.int64, device="cuda")
eng.dbg_gemm(m_tl, A, B, M, N, K, bias=bias, relu=True, aux_out=tl)
torch.cuda.synchronize()
t = tl.cpu().numpy().reshape(16, 8)
i = int(os.environ.get('H_TILE', '3'))
print(("planes " if PL else "fp32   ") + f"DPPO_H_DBG={os.environ.get('DPPO_H_DBG', '0'):>3s} M={M} N={N} K={K}: {us:7.1f} us | tile {i}: mainloop {t[i,1]-t[i,0]:6d}  acc {t[i,2]-t[i,1]:5d}  "
      f"tmem hold {t[i,3]-t[i,2]:6d}  write2 {t[i,4]-t[i,3]:6d}  period {t[i+1,0]-t[i,0]:6d} | mma thread waits: A {t[i,5]:6d} W {t[i,6]:6d} peer {t[i,7]:6d}", flush=True)
