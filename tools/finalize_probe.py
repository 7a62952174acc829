#!/usr/bin/env python3
"""Where the microseconds after the scan go: %globaltimer stamps inside finalize_kernel (world of one
rank, so the exchange is a loop-back) for k = 10 / 100, fp32 / bf16, host and device outputs."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_audio_search_b200 import SegmentIndex, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
names = ["select", "rescore", "push+flag", "wait", "merge+emit", "total"]
for dtype in ("fp32", "bf16"):
    idx = SegmentIndex(dtype, capacity=n)
    idx.append_synth(20261018, n, 0, n, n_queries=64, plants=200)
    idx.peer_attach(idx.peer_init(0, 1, max_queries=8, max_k=128))
    idx.set_option("stamp_exchange", 1)
    q = synth.raw_queries(20261018, 0, 64)
    qd = torch.from_numpy(q).cuda()
    for k in (10, 100):
        for to_host in (True, False):
            for i in range(40):
                idx.search_sharded(q[i:i + 1] if to_host else qd[i:i + 1], 0.5, 0.5, k=k, to_host=to_host)
            torch.cuda.synchronize()
            st = idx.exchange_stamps(32).astype(np.int64)
            d = np.stack([st[:, 1] - st[:, 0], st[:, 2] - st[:, 1], st[:, 3] - st[:, 2], st[:, 4] - st[:, 3], st[:, 5] - st[:, 4], st[:, 5] - st[:, 0]], 1) / 1e3
            t0 = time.perf_counter()
            for i in range(40):
                idx.search(q[i:i + 1], 0.5, 0.5, k=k)
            e2e = (time.perf_counter() - t0) / 40 * 1e3
            idx.set_option("time_kernels", 1)
            idx.search(q[:1], 0.5, 0.5, k=k)
            scan = idx.last_scan_ms()
            idx.set_option("time_kernels", 0)
            print(f"{dtype} k={k:3d} {'host' if to_host else 'dev '} outputs: " + "  ".join(f"{nm} {v:6.2f}" for nm, v in zip(names, d.mean(0))) +
                  f"  us | plain search e2e {e2e:.4f} ms, scan {scan:.4f} ms", flush=True)
    idx.close()
