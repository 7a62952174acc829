#!/usr/bin/env python3
"""GPU: sweep the GEMV scan configurations (unroll x blocks/SM x variant) and print scan-kernel
time + achieved algorithmic GB/s.  Usage: python tools/sweep_gemv.py [rows] [dtype ...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_audio_search_b200 import SegmentIndex, synth  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dtypes = sys.argv[2:] or ["fp32", "bf16"]
variants = [int(v) for v in os.environ.get("CAB_VARIANTS", "0").split(",")]
q = synth.raw_queries(1, 0, 64)
out = []
for dtype in dtypes:
    idx = SegmentIndex(dtype, capacity=rows)
    idx.append_synth(1, rows, 0, rows, n_queries=8, plants=30)
    idx.set_option("time_kernels", 1)
    bytes_ = rows * 2 * 384 * (4 if dtype == "fp32" else 2)
    for variant in variants:
        for unroll, bps in ((8, 1), (4, 2), (4, 1), (2, 4), (2, 3), (2, 2), (2, 1), (1, 4), (1, 3), (1, 2)):
            idx.set_option("gemv_variant", variant)
            idx.set_option("gemv_unroll", unroll)
            idx.set_option("gemv_blocks_per_sm", bps)
            ms = []
            for i in range(30):
                idx.search(q[i % 64], 0.5, 0.5, k=10)
                ms.append(idx.last_scan_ms())
            ms = np.array(ms[5:])
            rec = {"rows": rows, "dtype": dtype, "variant": variant, "unroll": unroll, "bps": bps,
                   "scan_ms_mean": float(ms.mean()), "scan_ms_min": float(ms.min()),
                   "gbs_mean": bytes_ / ms.mean() / 1e6, "gbs_best": bytes_ / ms.min() / 1e6}
            out.append(rec)
            print(json.dumps(rec), flush=True)
    idx.close()

# multi-query register tiling: queries per second for a batch of 32 queries
for dtype in dtypes:
    idx = SegmentIndex(dtype, capacity=rows)
    idx.append_synth(1, rows, 0, rows, n_queries=8, plants=30)
    idx.set_option("time_kernels", 1)
    for tile, unroll in ((1, 0), (2, 0), (4, 0), (4, 4)):
        if dtype == "bf16" and unroll == 4:
            continue
        idx.set_option("gemv_query_tile", tile)
        idx.set_option("gemv_unroll", unroll)
        ms = []
        for i in range(8):
            idx.search(q[:32], 0.5, 0.5, k=10, path="gemv")
            ms.append(idx.last_scan_ms())
        ms = float(np.mean(ms[2:]))
        print(json.dumps({"rows": rows, "dtype": dtype, "query_tile": tile, "unroll": unroll, "batch": 32, "scan_ms_total": ms,
                          "queries_per_s": 32e3 / ms, "corpus_passes": 32 // tile}), flush=True)
    idx.close()
