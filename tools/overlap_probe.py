#!/usr/bin/env python3
"""Upper bound for scan/scan overlap: two handles (own workspaces, own copies of the 1M library) searched
alternately on two streams, so the tail of one scan overlaps the head of the next."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_audio_search_b200 import SegmentIndex, synth
n = 1_000_000
idx = [SegmentIndex("fp32", capacity=n) for _ in range(2)]
for i in idx:
    i.append_synth(20261018, n, 0, n, n_queries=64, plants=30); i.set_option("queries_settled", 1)
q = torch.from_numpy(synth.raw_queries(20261018, 0, 64)).cuda()
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
def run(two, steps=200):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    for i in range(steps):
        h = i & 1 if two else 0
        with torch.cuda.stream(streams[h]):
            idx[h].search(q[i % 64:i % 64 + 1], 0.5, 0.5, k=10)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
for _ in range(3):
    print("one handle, one stream: %.4f ms/query   two handles, two streams: %.4f ms/query" % (run(False), run(True)), flush=True)
