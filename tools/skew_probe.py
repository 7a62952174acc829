#!/usr/bin/env python3
"""torchrun, N GPUs: where the sharded step's time over N=1 goes -- per-rank scan time (library
events around the scan kernel) against the lock-step step time of the p2p exchange."""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_audio_search_b200 import SegmentIndex, ShardedSearcher, shard_range, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
per = 1_000_000
n = per * world
lo, hi = shard_range(n, rank, world)
idx = SegmentIndex("fp32", capacity=hi - lo, device=local)
idx.append_synth(3, n, lo, hi, n_queries=8, plants=12, partial=False)
idx.row_base = lo
q = torch.from_numpy(synth.raw_queries(3, 0, 8)).cuda()
sh = ShardedSearcher(idx, rank, world, exchange="p2p", max_queries=1, max_k=10)
steps = 300


def loop(fn, timed_scan=False):
    for i in range(20):
        fn(i)
    dist.barrier(); torch.cuda.synchronize()
    scans = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
        if timed_scan:
            scans.append(idx.last_scan_ms())
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (float(np.mean(scans)) if scans else None)


local_ms, _ = loop(lambda i: idx.search(q[i % 8:i % 8 + 1], 0.5, 0.5, k=10))                 # this shard alone, no exchange
shard_ms, _ = loop(lambda i: sh.search(q[i % 8:i % 8 + 1], 0.5, 0.5, k=10, to_host=False))  # lock-step with the peers
idx.set_option("time_kernels", 1)
_, scan_ms = loop(lambda i: idx.search(q[i % 8:i % 8 + 1], 0.5, 0.5, k=10), timed_scan=True)
rows = [None] * world
dist.all_gather_object(rows, {"rank": rank, "local_step_ms": round(local_ms, 4), "sharded_step_ms": round(shard_ms, 4),
                              "scan_ms": round(scan_ms, 4)})
if rank == 0:
    for r in rows:
        print(json.dumps(r))
    slow = max(r["local_step_ms"] for r in rows)
    print(json.dumps({"world": world, "slowest_local_step_ms": slow, "sharded_step_ms": max(r["sharded_step_ms"] for r in rows),
                      "exchange_and_lockstep_overhead_us": round(1e3 * (max(r["sharded_step_ms"] for r in rows) - slow), 1)}))
dist.destroy_process_group()
