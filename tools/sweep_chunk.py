import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_audio_search_b200 import SegmentIndex, synth
q = synth.raw_queries(1, 0, 64)
for rows, dtype in ((1_000_000, "fp32"), (1_000_000, "bf16"), (10_000_000, "fp32")):
    idx = SegmentIndex(dtype, capacity=rows); idx.append_synth(1, rows, 0, rows, n_queries=8, plants=30)
    idx.set_option("time_kernels", 1)
    for chunk in (8, 16, 32, 64, 128, 256):
        idx.set_option("gemv_chunk_rows", chunk)
        ms = []
        for i in range(40):
            idx.search(q[i % 64], 0.5, 0.5, k=10); ms.append(idx.last_scan_ms())
        ms = np.array(ms[8:])
        print(rows, dtype, "chunk", chunk, "scan ms", round(float(ms.mean()), 4), "GB/s", round(rows * 768 * (4 if dtype == "fp32" else 2) / ms.mean() / 1e6), flush=True)
    idx.close()
