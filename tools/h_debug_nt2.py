#!/usr/bin/env python
"""Bad output rows of the dX epilogue test (mode 97): does the row look like the product with a STALE k-block of the activation
operand (the raw-ring slot's previous occupant)?  Prints one line per hypothesis that explains a row."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tensorflow-dppo_b200"))
import numpy as np, torch
from dppo_b200 import DppoEngine
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (20000, 256, 256)
rng = np.random.default_rng(5)
eng = DppoEngine(376, 17, (256, 256), max(M, 32768), gemm_mode=1)
A = rng.standard_normal((M, K)).astype(np.float32)
Bt = (rng.standard_normal((N, K)) / math.sqrt(K)).astype(np.float32)
mask = rng.standard_normal((M, N)).astype(np.float32)
A64, B64 = A.astype(np.float64), Bt.astype(np.float64)
ref = A64 @ B64.T
ref2 = ref * (mask > 0)
dA, dB, dM = (torch.from_numpy(x).cuda() for x in (A, Bt, mask))
nkb = (K + 31) // 32
for trial in range(2):
    got = eng.dbg_gemm(97, dA, dB, M, N, K, mask=dM).cpu().numpy().astype(np.float64)
    bad = np.abs(got - ref2) > 1e-4
    rows = np.unique(np.nonzero(bad)[0])
    print(f"trial {trial}: bad elements {bad.sum()}, bad rows {rows.size}, tiles {np.unique(rows // 128)[:20].tolist()}")
    # do the two column halves (different epilogue warps) share the bad rows?
    left = np.unique(np.nonzero(bad[:, :32])[0]); right = np.unique(np.nonzero(bad[:, 32:64])[0])
    print(f"  rows bad in columns 0-31: {left.size}, in 32-63: {right.size}, in both: {np.intersect1d(left, right).size}")
    explained = 0
    for r in rows[:40]:
        keep = mask[r] > 0
        hit = None
        for kb in range(nkb):
            ks = slice(kb * 32, min(K, kb * 32 + 32))
            base = ref[r] - A64[r, ks] @ B64[:, ks].T
            cands = []
            for d in (-4, -3, -2, -1, 1, 2, 3, 4):                      # same row, another k-block
                if 0 <= kb + d < nkb and (kb + d) * 32 + 32 <= K and ks.stop - ks.start == 32:
                    cands.append((f"same row, k-block {kb + d} instead of {kb}", A64[r, (kb + d) * 32:(kb + d) * 32 + 32]))
            for dr in (-256, -128, 128, 256):                            # another tile's row, any k-block
                if 0 <= r + dr < M:
                    for kb2 in range(nkb):
                        if kb2 * 32 + 32 <= K and ks.stop - ks.start == 32:
                            cands.append((f"row {dr:+d}, k-block {kb2} instead of {kb}", A64[r + dr, kb2 * 32:kb2 * 32 + 32]))
            cands.append((f"zeros instead of k-block {kb}", np.zeros(ks.stop - ks.start)))
            for name, vec in cands:
                pred = (base + vec @ B64[:, ks].T) * keep
                if np.abs(pred - got[r]).max() < 1e-4:
                    hit = name; break
            if hit: break
        if hit:
            explained += 1
            if explained <= 12: print(f"  row {r} (tile {r // 128}, row {r % 128}): {hit}")
    print(f"  explained {explained} of {min(40, rows.size)} rows by ONE substituted k-block")
