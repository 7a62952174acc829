#!/usr/bin/env python3
"""Times cab_score_all (legacy all-N scoring, previous_iterations/streamlit_app.py:173-223) on one
GPU: device-resident query and output (kernel only) and the host call (query in the kernel
arguments, N fp32 scores copied back).  One JSON line per configuration.

    python tools/bench_score_all.py [steps] [--no-flush] [--only N:dtype]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_audio_search_b200 import SegmentIndex, synth  # noqa: E402

ADAPTIVE = ((0.2, 0.8), (0.7, 0.3), (0.2, 0.8), (0.2, 0.8))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    steps = int(args[0]) if args else 50
    do_flush = "--no-flush" not in sys.argv
    configs = ((1_000_000, "fp32"), (10_000_000, "fp32"), (10_000_000, "bf16"))
    if "--only" in sys.argv:
        n_s, dt = sys.argv[sys.argv.index("--only") + 1].split(":")
        configs = ((int(n_s), dt),)
    peak = 6549.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    for n, dtype in configs:
        idx = SegmentIndex(dtype, capacity=n, device=0)
        idx.append_synth(7, n, 0, n, n_queries=4, plants=8, partial=False)
        q = synth.raw_queries(7, 0, 4)
        qd = torch.from_numpy(q).cuda()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
        # CUDA events recorded by the library right around the launch, on the launching stream
        # (events recorded from Python would also count the host time between the two records)
        idx.set_option("time_kernels", 1)
        for i in range(5):
            idx.score_all(qd[i % 4:i % 4 + 1], ADAPTIVE)
        torch.cuda.synchronize()
        times = []
        for i in range(steps):
            if do_flush:
                flush.zero_()
            idx.score_all(qd[i % 4:i % 4 + 1], ADAPTIVE)
            times.append(idx.last_scan_ms())
        ms = float(np.median(times))
        idx.set_option("time_kernels", 0)
        row_bytes = 2 * 384 * (4 if dtype == "fp32" else 2) + 4            # both corpora read + one fp32 score written
        gbs = n * row_bytes / ms / 1e6
        t0 = time.perf_counter()
        for i in range(10):
            idx.score_all(q[i % 4], ADAPTIVE)
        host_ms = (time.perf_counter() - t0) / 10 * 1e3
        pinned = idx.pinned_scores(1)
        idx.score_all(q[0], ADAPTIVE, out=pinned)
        t0 = time.perf_counter()
        for i in range(10):
            idx.score_all(q[i % 4], ADAPTIVE, out=pinned)
        pinned_ms = (time.perf_counter() - t0) / 10 * 1e3
        print(json.dumps({"kernel": "score_all", "segments": n, "dtype": dtype, "steps": steps,
                          "kernel_ms": round(ms, 4), "algorithmic_GBps": round(gbs, 1),
                          "frac_of_measured_peak": round(gbs / peak, 3), "frac_of_8TBps": round(gbs / 8000, 3),
                          "host_call_ms": round(host_ms, 3), "host_call_pinned_out_ms": round(pinned_ms, 3), "d2h_bytes": 4 * n,
                          "l2": "flushed between steps (256 MB memset)" if do_flush else "not flushed (corpus >> 126 MB L2)"}), flush=True)
        idx.close()
        del flush
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
