import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_audio_search_b200 import SegmentIndex, synth
rows, nq, k = int(sys.argv[1]), int(sys.argv[2]), 10
seed = 20261018
q = synth.raw_queries(seed, 0, nq)
W = [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]
wa = np.array([W[i % 7] for i in range(nq)]); wb = 1 - wa
bf = SegmentIndex("bf16", capacity=rows)
bf.append_synth(seed, rows, 0, rows, n_queries=nq, plants=20)
ref = bf.search(q, wa, wb, k=k, path="gemm")
qd = torch.from_numpy(q).cuda()
for path in ("gemv", "gemm"):
    g = bf.search(qd[:64] if path == "gemv" else qd, wa[:64] if path == "gemv" else wa, wb[:64] if path == "gemv" else wb, k=k, path=path)
    torch.cuda.synchronize()
    gi = g.indices.cpu().numpy(); gf = g.fusion.cpu().numpy(); gc = g.count.cpu().numpy()
    n = gi.shape[0]
    print(path, "identical", int((gi == ref.indices[:n]).all(axis=1).sum()), "of", n)
    for r in (0, 1, 5):
        print("  q", r, "dev", gi[r][:5], np.round(gf[r][:5], 4), gc[r], "| ref", ref.indices[r][:5], np.round(ref.fusion[r][:5], 4), ref.count[r])
if nq > 256:
    print("per-pass identical (gemm device):", [int((gi[p:p+256] == ref.indices[p:p+256]).all(axis=1).sum()) for p in range(0, nq, 256)])
    for r in (256, 300, 700):
        print("  q", r, "dev", gi[r][:5], np.round(gf[r][:5], 4), gc[r], "| ref", ref.indices[r][:5], np.round(ref.fusion[r][:5], 4), ref.count[r])
    # second device call
    g2 = bf.search(qd, wa, wb, k=k, path="gemm"); torch.cuda.synchronize()
    gi2 = g2.indices.cpu().numpy()
    print("second call per-pass identical:", [int((gi2[p:p+256] == ref.indices[p:p+256]).all(axis=1).sum()) for p in range(0, nq, 256)])
