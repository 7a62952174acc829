#!/usr/bin/env python3
"""Debug aid: 2 processes on cuda:0 (gloo for the handle exchange), sharded multi-batch k=100 vs one index vs the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(rank, world, port, two_gpus):
    import torch, torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimodal_audio_search_b200 import SegmentIndex, ShardedSearcher, synth
    from oracle import numpy_oracle as no
    dev = rank if two_gpus else 0
    torch.cuda.set_device(dev)
    SEED, per, nq = 20261019, 60_000, 40
    n_total = per * world
    whole = SegmentIndex("fp32", capacity=n_total, device=dev)
    whole.append_synth(SEED, n_total, 0, n_total, n_queries=nq, plants=30, partial=True)
    part = SegmentIndex("fp32", capacity=per, device=dev)
    part.append_synth(SEED, n_total, rank * per, (rank + 1) * per, n_queries=nq, plants=30, partial=True)
    part.row_base = rank * per
    sh = ShardedSearcher(part, rank, world, exchange="p2p", max_queries=64, max_k=100)
    q = synth.raw_queries(SEED, 0, nq)
    wa = np.array([[0.5, 0.2, 0.3, 0.4, 0.6, 0.7, 0.8][i % 7] for i in range(nq)]); wb = 1 - wa
    a, b, f, _ = synth.library(SEED, n_total, nq, 30, True)
    qd = torch.from_numpy(q).cuda()
    seq = [(i, i + 1, 10) for i in range(8)] + [(8, 9, 100), (0, 40, 10), (9, 10, 10), (0, 40, 100)] if os.environ.get("BENCH_SEQ") else []
    for lo, hi, kk in seq + [(0, 40, 100), (0, 33, 100), (0, 32, 100), (0, 40, 50), (0, 40, 64), (0, 40, 65), (0, 40, 10), (0, 40, 100)]:
        want = whole.search(q[lo:hi], wa[lo:hi], wb[lo:hi], k=kk)
        gotd = sh.search(qd[lo:hi], wa[lo:hi], wb[lo:hi], k=kk, to_host=False)
        got = sh.search(q[lo:hi], wa[lo:hi], wb[lo:hi], k=kk, to_host=True)
        if not np.array_equal(gotd.indices.cpu().numpy(), got.indices):
            print(f"rank {rank} case {(lo, hi, kk)}: device variant != host variant", flush=True)
        nccl_like = part.search_candidates(q[lo:hi], wa[lo:hi], wb[lo:hi], k=kk)
        torch.cuda.synchronize()
        bad_w = bad_g = 0
        for j in range(hi - lo):
            o = no.search(q[lo + j], a, b, f, wa[lo + j], wb[lo + j], k=kk)
            bad_w += list(want.indices[j, :len(o.indices)]) != list(o.indices)
            bad_g += list(got.indices[j, :len(o.indices)]) != list(o.indices)
        if os.environ.get("SNAP") and bad_g:
            # what does the exchange buffer hold?  expected: every rank's search_candidates block
            mine = nccl_like.cpu().numpy().copy()                      # [Q, k, 24] u8
            blocks = [None] * world
            dist.all_gather_object(blocks, mine)
            snap = part.peer_snapshot()
            flags = snap[:16].view(np.uint32)
            half = world * 64 * 100
            body = snap[256:].reshape(2, half, 24)
            nqc = hi - lo
            for par in (0, 1):
                for r in range(world):
                    seg = body[par, r * nqc * kk:(r + 1) * nqc * kk].reshape(nqc, kk, 24)
                    same = (seg == blocks[r]).all(axis=(1, 2))
                    idxs = seg.view(np.int64)[..., 0] if False else np.ascontiguousarray(seg[..., :8]).view(np.int64)[..., 0]
                    print(f"  rank {rank} snapshot parity {par} list {r}: queries equal to rank {r}'s candidates: {int(same.sum())}/{nqc}; "
                          f"first mismatching query {int(np.argmin(same)) if not same.all() else -1}; flags {flags.tolist()}; "
                          f"q0 first idx {idxs[0, :3].tolist()} expected {np.ascontiguousarray(blocks[r][0, :3, :8]).view(np.int64)[:, 0].tolist()}", flush=True)
        cand = nccl_like.cpu().numpy().view(np.dtype([("index", "<i8"), ("a", "<f4"), ("b", "<f4"), ("fl", "<u4"), ("pad", "<u4")]))
        print(f"rank {rank} case {(lo, hi, kk)}: whole!=oracle {bad_w}, sharded!=oracle {bad_g}; "
              f"q0 sharded rows>=per: {(got.indices[0] >= per).sum()} want: {(want.indices[0] >= per).sum()}; "
              f"local cand q0 valid {(cand['index'][0, :, 0] >= 0).sum()}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    mp.spawn(worker, args=(2, 29711, "--two-gpus" in sys.argv), nprocs=2)
