#!/usr/bin/env python3
"""One GPU: what the sharded step's structure costs without any NVLink traffic -- fused
scan+finalize/emit (2 launches) vs scan + finalize + separate emit (3 launches, the N>1 shape)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_audio_search_b200 import SegmentIndex, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
idx = SegmentIndex("fp32", capacity=n, device=0)
idx.append_synth(3, n, 0, n, n_queries=8, plants=12, partial=False)
q = torch.from_numpy(synth.raw_queries(3, 0, 8)).cuda()
steps = 300


def timed(fn):
    for i in range(20):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def fused(i):
    idx.search(q[i % 8:i % 8 + 1], 0.5, 0.5, k=10)


def unfused(i):
    c = idx.search_candidates(q[i % 8:i % 8 + 1], 0.5, 0.5, k=10)
    idx.merge_candidates(c.unsqueeze(0), None, None, k=10, to_host=False)


print(json.dumps({"segments": n, "fused_ms": round(timed(fused), 4), "unfused_ms": round(timed(unfused), 4)}))
