#!/usr/bin/env python3
"""Where the tensor-core scan overtakes the register-tiled GEMV: ms per call for nq queries, both paths."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_audio_search_b200 import SegmentIndex, synth

def t(idx, q, path, k=10, reps=8):
    wa = np.full(q.shape[0], 0.5); wb = 1 - wa
    for _ in range(2): idx.search(q, wa, wb, k=k, path=path)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): idx.search(q, wa, wb, k=k, path=path)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for n in (100_000, 1_000_000, 10_000_000):
    for dtype in ("bf16", "fp32"):
        idx = SegmentIndex(dtype, capacity=n)
        idx.append_synth(20261018, n, 0, n, n_queries=64, plants=30)
        if dtype == "fp32":
            idx.enable_tensor_core_batches()
        qs = torch.from_numpy(synth.raw_queries(20261018, 0, 64)).cuda()
        row = []
        for nq in (2, 4, 8, 16, 32, 64):
            row.append(f"nq={nq}: gemv {t(idx, qs[:nq], 'gemv'):.3f} / tc {t(idx, qs[:nq], 'gemm'):.3f}")
        print(f"{n:>9} {dtype}: " + " | ".join(row), flush=True)
        idx.close()
