#!/usr/bin/env python3
"""BASELINE.json config 5 (GPU): keyword-weight sweep (20 %-80 % ASR weight) over a 10 M-segment
library, 4096 mixed queries; recall@10 of the bf16 tensor-core path against the fp32 engine's
ordering (the fp32 engine itself is parity-tested against the oracle/reference goldens).

    python tools/recall_sweep.py [rows] [queries]   -> JSON on stdout + profiles-style markdown
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from multimodal_audio_search_b200 import SegmentIndex, synth  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
seed, k, plants = 20261018, 10, 20
W = [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]          # the achievable weight classes

q = synth.raw_queries(seed, 0, nq)
wa = np.array([W[i % len(W)] for i in range(nq)])
wb = 1.0 - wa
out = {"rows": rows, "queries": nq, "k": k, "plants_per_query": plants}

f32 = SegmentIndex("fp32", capacity=rows)
f32.append_synth(seed, rows, 0, rows, n_queries=nq, plants=plants)
bf = SegmentIndex("bf16", capacity=rows)
bf.append_synth(seed, rows, 0, rows, n_queries=nq, plants=plants)
qd = torch.from_numpy(q).cuda()

torch.cuda.synchronize(); t0 = time.perf_counter()
ref_idx, ref_f = [], []
for i in range(0, nq, 32):
    r = f32.search(qd[i:i + 32], wa[i:i + 32], wb[i:i + 32], k=k, path="gemv")
    ref_idx.append(r.indices.cpu().numpy()); ref_f.append(r.fusion.cpu().numpy())
torch.cuda.synchronize(); out["fp32_gemv_seconds"] = time.perf_counter() - t0
ref_idx, ref_f = np.concatenate(ref_idx), np.concatenate(ref_f)

res = {}
for path in ("gemm", "gemv"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if path == "gemm":
        r = bf.search(qd, wa, wb, k=k, path="gemm")
        gi, gf = r.indices.cpu().numpy(), r.fusion.cpu().numpy()
    else:
        gi, gf = [], []
        for i in range(0, nq, 32):
            r = bf.search(qd[i:i + 32], wa[i:i + 32], wb[i:i + 32], k=k, path="gemv")
            gi.append(r.indices.cpu().numpy()); gf.append(r.fusion.cpu().numpy())
        gi, gf = np.concatenate(gi), np.concatenate(gf)
    torch.cuda.synchronize(); secs = time.perf_counter() - t0
    hits = np.array([len(set(gi[i][gi[i] >= 0]) & set(ref_idx[i][ref_idx[i] >= 0])) for i in range(nq)])
    tot = np.array([(ref_idx[i] >= 0).sum() for i in range(nq)])
    # score error on the common ranks
    both = (gi == ref_idx) & (ref_idx >= 0)
    err = float(np.abs(gf - ref_f)[both].max()) if both.any() else 0.0
    per_w = {str(w): float(hits[np.isclose(wa, w)].sum() / max(1, tot[np.isclose(wa, w)].sum())) for w in W}
    res[path] = {"recall_at_10": float(hits.sum() / tot.sum()), "per_asr_weight": per_w,
                 "max_abs_score_error_same_rank": err, "seconds": secs,
                 "identical_lists": int((gi == ref_idx).all(axis=1).sum())}
out["bf16_vs_fp32"] = res
print(json.dumps(out))
