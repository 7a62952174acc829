#!/usr/bin/env python3
"""Ingest-side timings: append from host numpy, append from CUDA
tensors, index save / load, and the columnar segment table.  One JSON line each.

    python tools/bench_ingest.py [rows]
"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_audio_search_b200 import SegmentIndex, SegmentTable  # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    rng = np.random.default_rng(0)
    block = rng.standard_normal((65536, 384), dtype=np.float32)
    a = np.tile(block, (-(-n // 65536), 1))[:n]
    b = a[::-1].copy()
    flags = rng.integers(1, 4, n).astype(np.uint8)
    raw_gb = 2 * n * 384 * 4 / 1e9
    for dtype in ("fp32", "bf16"):
        idx = SegmentIndex(dtype, capacity=n, device=0)
        idx.append(a[:70000], b[:70000], flags[:70000])             # warm-up: staging buffers
        idx.clear()
        t0 = time.perf_counter()
        idx.append(a, b, flags)
        dt = time.perf_counter() - t0
        emit(op="append_host", dtype=dtype, rows=n, seconds=round(dt, 4), rows_per_s=round(n / dt),
             host_GBps=round(raw_gb / dt, 2), note="pageable numpy -> pinned slots -> H2D -> normalise kernels, capacity reserved")
        idx.close()
        idx = SegmentIndex(dtype, capacity=0, device=0)
        t0 = time.perf_counter()
        for r in range(0, n, 50000):                                  # incremental: geometric growth
            idx.append(a[r:r + 50000], b[r:r + 50000], flags[r:r + 50000])
        dt = time.perf_counter() - t0
        emit(op="append_host_incremental_50k", dtype=dtype, rows=n, seconds=round(dt, 4), rows_per_s=round(n / dt),
             note="no reserved capacity: geometric growth with device-to-device copies")
        idx.close()
        da, db, df = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(flags).cuda()
        idx = SegmentIndex(dtype, capacity=n, device=0)
        idx.append(da[:1000], db[:1000], df[:1000]); idx.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        idx.append(da, db, df)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out_gb = 2 * n * 384 * (4 if dtype == "fp32" else 2) / 1e9
        emit(op="append_device", dtype=dtype, rows=n, seconds=round(dt, 5), rows_per_s=round(n / dt),
             hbm_GBps=round((raw_gb + out_gb) / dt, 1), note="CUDA tensors: normalise kernels only (read raw fp32, write rows)")
        del da, db, df
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
            path = os.path.join(d, "lib.cab")
            t0 = time.perf_counter()
            idx.save(path)
            dt_s = time.perf_counter() - t0
            size = os.path.getsize(path)
            idx.close()
            t0 = time.perf_counter()
            idx = SegmentIndex.load(path, device=0)
            dt_l = time.perf_counter() - t0
            emit(op="index_file", dtype=dtype, rows=n, file_GB=round(size / 1e9, 3), save_s=round(dt_s, 3),
                 load_s=round(dt_l, 3), save_GBps=round(size / 1e9 / dt_s, 2), load_GBps=round(size / 1e9 / dt_l, 2),
                 where=d.split("/")[1])
        idx.close()
        torch.cuda.empty_cache()
    # columnar records
    m = min(n, 200_000)
    segs = [{"segment_id": f"seg_{i}", "start_time": 5.0 * i, "end_time": 5.0 * i + 10, "duration": 10.0,
             "asr_text": "some words that were said", "asr_embedding": block[i % 65536], "asr_success": True,
             "audio_description": "a sound", "audio_embedding": block[(i + 1) % 65536], "audio_success": True,
             "audio_data": None, "sample_rate": 16000} for i in range(m)]
    t0 = time.perf_counter()
    t = SegmentTable.from_segments(segs)
    dt_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    t.drain_pending()
    dt_drain = time.perf_counter() - t0
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "lib.meta")
        t0 = time.perf_counter()
        t.save(path)
        dt_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        u = SegmentTable.load(path)
        recs = [u[i] for i in range(0, m, m // 10)][:10]
        dt_l = time.perf_counter() - t0
    emit(op="segment_table", rows=m, extend_s=round(dt_build, 3), rows_per_s=round(m / dt_build),
         drain_embeddings_s=round(dt_drain, 3), save_s=round(dt_s, 3), load_plus_10_records_ms=round(dt_l * 1e3, 2),
         n_records=len(recs))


if __name__ == "__main__":
    main()
