"""torchrun, N GPUs: the peer-memory (p2p) sharded search against the NCCL all-gather path and
against a single index holding the whole library on rank 0."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from multimodal_audio_search_b200 import SegmentIndex, ShardedSearcher, shard_range, synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
seed, n = 99, 300_001
for dtype, nq, k, path in (("fp32", 1, 10, "gemv"), ("fp32", 3, 10, "gemv"), ("bf16", 70, 100, "gemm"), ("fp32", 40, 37, "gemv"), ("bf16", 1, 100, "gemv")):
    lo, hi = shard_range(n, rank, world)
    idx = SegmentIndex(dtype, capacity=hi - lo, device=local)
    idx.append_synth(seed, n, lo, hi, n_queries=nq, plants=40, partial=True)
    idx.row_base = lo
    q = synth.raw_queries(seed, 0, nq)
    wa = np.linspace(0.2, 0.8, nq); wb = 1 - wa
    p2p = ShardedSearcher(idx, rank, world, exchange="p2p", max_queries=nq, max_k=k)
    nccl = ShardedSearcher(idx, rank, world, exchange="nccl")
    for rep in range(5):                                   # several epochs: both buffer parities, flag reuse
        a = p2p.search(q, wa, wb, k=k, path=path)
        b = nccl.search(q, wa, wb, k=k, path=path)
        assert a.indices.tolist() == b.indices.tolist() and a.fusion.tolist() == b.fusion.tolist() and a.count.tolist() == b.count.tolist(), (dtype, rep)
    d = p2p.search(torch.from_numpy(q).cuda(), wa, wb, k=k, path=path, to_host=False)
    assert d.indices.cpu().numpy().tolist() == a.indices.tolist()
    if rank == 0:
        whole = SegmentIndex(dtype, capacity=n, device=local)
        whole.append_synth(seed, n, 0, n, n_queries=nq, plants=40, partial=True)
        w = whole.search(q, wa, wb, k=k, path=path)
        same = (w.indices == a.indices).all(axis=1).mean()
        assert same >= (1.0 if path == "gemv" else 0.6), same      # gemm: per-shard bf16-query selection may differ at the k-th boundary
        print(f"{dtype} Q={nq} k={k} {path}: p2p == nccl on all ranks; identical to the unsharded index for {same*100:.0f}% of queries", flush=True)
    dist.barrier()
# legacy all-N scoring over the shards: one all-gather of the score vectors, every rank gets [Q, n]
for dtype in ("fp32", "bf16"):
    lo, hi = shard_range(n, rank, world)
    idx = SegmentIndex(dtype, capacity=hi - lo, device=local)
    idx.append_synth(seed, n, lo, hi, n_queries=2, plants=40, partial=True)
    idx.row_base = lo
    cls = (np.arange(lo, hi) % 3).astype(np.uint8)
    idx.set_weight_classes(cls)
    q = torch.from_numpy(synth.raw_queries(seed, 0, 2)).cuda()
    table = [[0.2, 0.8], [0.7, 0.3], [1.0, 0.0], [0.0, 1.0]]
    full = ShardedSearcher(idx, rank, world).score_all(q, table, n)
    assert full.shape == (2, n) and full.is_cuda
    if rank == 0:
        whole = SegmentIndex(dtype, capacity=n, device=local)
        whole.append_synth(seed, n, 0, n, n_queries=2, plants=40, partial=True)
        whole.set_weight_classes((np.arange(n) % 3).astype(np.uint8))
        ref = whole.score_all(q, table)
        assert torch.equal(full, ref), dtype
        print(f"{dtype}: sharded score_all == unsharded, bit for bit ({world} ranks, {n} segments)", flush=True)
    dist.barrier()
if rank == 0: print("p2p check ok")
dist.destroy_process_group()
