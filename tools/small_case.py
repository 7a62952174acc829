"""Small end-to-end case for compute-sanitizer: both dtypes, GEMV + tensor-core paths, finalize
fast/general paths, sharded merge, save/load."""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_audio_search_b200 import SegmentIndex, synth
seed, n = 3, 3000
q = synth.raw_queries(seed, 0, 70)
wa = np.linspace(0.2, 0.8, 70); wb = 1 - wa
for dtype in ("fp32", "bf16"):
    idx = SegmentIndex(dtype)
    idx.append_synth(seed, n, 0, n - 7, n_queries=8, plants=20, partial=True)
    a, b, f, _ = synth.library(seed, n, 8, 20, True, r0=n - 7, r1=n)
    idx.append(a, b, f)
    r1 = idx.search(q[:3], wa[:3], wb[:3], k=10)
    idx.set_option("finalize_general", 1)
    r2 = idx.search(q[:3], wa[:3], wb[:3], k=10)
    idx.set_option("finalize_general", 0)
    assert r1.indices.tolist() == r2.indices.tolist()
    r3 = idx.search(q[:5], wa[:5], wb[:5], k=128, threshold=-1.0)
    if dtype == "bf16":
        g = idx.search(q, wa, wb, k=100, path="gemm", threshold=-1.0)
        assert (g.count == 100).all()
        d = idx.search(torch.from_numpy(q).cuda(), wa, wb, k=10, path="gemm")
        torch.cuda.synchronize()
    c = idx.search_candidates(q[:2], wa[:2], wb[:2], k=10)
    m = idx.merge_candidates(torch.stack([c, c]).contiguous(), wa[:2], wb[:2], k=10)
    with tempfile.TemporaryDirectory() as t:
        idx.save(os.path.join(t, "x.cab"))
        back = SegmentIndex.load(os.path.join(t, "x.cab"), rows=(10, 2000))
        back.search(q[:1], 0.5, 0.5)
    idx.close()
print("sanitize case ok")
