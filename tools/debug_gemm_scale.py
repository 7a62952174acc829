import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_audio_search_b200 import SegmentIndex, synth
seed = 20261018
NQ = 4096
q = synth.raw_queries(seed, 0, NQ)
W = [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]
wa = np.array([W[i % 7] for i in range(NQ)]); wb = 1 - wa
rows = int(sys.argv[1])
bf = SegmentIndex("bf16", capacity=rows)
bf.append_synth(seed, rows, 0, rows, n_queries=4096, plants=20)
qd = torch.from_numpy(q).cuda()
k = 10
ref = bf.search(q[:512], wa[:512], wb[:512], k=k, path="gemm")       # host path, 2 passes
for nq in (256, 512, 1024, 2048, 4096):
    for mode in ("host", "device"):
        g = bf.search(q[:nq] if mode == "host" else qd[:nq], wa[:nq], wb[:nq], k=k, path="gemm")
        gi = g.indices if mode == "host" else g.indices.cpu().numpy()
        gc = g.count if mode == "host" else g.count.cpu().numpy()
        print(f"rows {rows} nq {nq} {mode}: first-512 identical {(gi[:512] == ref.indices[:min(512,nq)]).all(axis=1).sum()} counts {gc[:4]} {gc[-4:]} zero-count {int((gc==0).sum())}", flush=True)
