#!/usr/bin/env python3
"""Throughput of concurrent sessions through the SearchBatcher vs one query per call.

    python tools/bench_batcher.py [segments] [dtype] [threads] [queries_per_thread] [shadow]

`shadow` (fp32 only): keep bf16 shadows so that batches are preselected on the tensor cores.
"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_audio_search_b200 import SearchBatcher, SegmentIndex, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    dtype = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    n_threads = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    per_thread = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    shadow = len(sys.argv) > 5 and sys.argv[5] == "shadow" and dtype == "fp32"
    idx = SegmentIndex(dtype, capacity=n, device=0)
    if shadow:
        idx.enable_tensor_core_batches()
    idx.append_synth(11, n, 0, n, n_queries=64, plants=12, partial=False)
    q = synth.raw_queries(11, 0, 64)
    for i in range(5):
        idx.search(q[i:i + 1], 0.5, 0.5)
    total = n_threads * per_thread
    t0 = time.perf_counter()
    for i in range(total // 4):
        idx.search(q[i % 64:i % 64 + 1], 0.4, 0.6)
    serial_qps = (total // 4) / (time.perf_counter() - t0)
    for max_wait in (0.0, 0.0005):
        batcher = SearchBatcher(idx, max_batch=256, max_wait_s=max_wait)

        def session(t):
            for j in range(per_thread):
                batcher.search(q[(t + j) % 64], 0.4, 0.6)
        threads = [threading.Thread(target=session, args=(t,)) for t in range(n_threads)]
        t0 = time.perf_counter()
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        dt = time.perf_counter() - t0
        st = batcher.stats
        batcher.close()
        print(json.dumps({"segments": n, "dtype": dtype + (" + bf16 shadows" if shadow else ""), "threads": n_threads, "queries": total,
                          "max_wait_s": max_wait, "serial_queries_per_s": round(serial_qps, 1),
                          "batched_queries_per_s": round(total / dt, 1), "speedup": round(total / dt / serial_qps, 2),
                          "gpu_calls": st.batches, "mean_batch": round(st.mean_batch, 1), "largest_batch": st.largest_batch}),
              flush=True)
    idx.close()


if __name__ == "__main__":
    main()
