"""CPU, world_size 2 over gloo: the sharded-search plumbing (shard ranges, candidate packing, the
single all-gather, merge call order).  The two CUDA stages are replaced by oracle-backed fakes,
so this checks the host logic of multimodal_audio_search_b200/sharded.py without a GPU."""
import os
import socket

import numpy as np
import pytest

from multimodal_audio_search_b200 import synth
from multimodal_audio_search_b200.sharded import CANDIDATE_DTYPE, ShardedSearcher, shard_range
from oracle import numpy_oracle as no


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 1000, 100_000_001):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans if hi > lo or n == 0) <= -(-n // world)


class FakeShardIndex:
    """search_candidates / merge_candidates computed by the numpy oracle (test infrastructure)."""

    def __init__(self, seed, n_total, lo, hi, plants, nq):
        self.a, self.b, self.f, _ = synth.library(seed, n_total, nq, plants, True, r0=lo, r1=hi)
        self.lo = lo

    def search_candidates(self, queries, w_asr, w_audio, k, threshold, path):
        import torch
        self._w = (w_asr, w_audio)
        out = np.zeros((len(queries), k), dtype=CANDIDATE_DTYPE)
        out["index"] = -1
        for i, q in enumerate(queries):
            o = no.search(q, self.a, self.b, self.f, w_asr[i], w_audio[i], k=k, threshold=threshold)
            n = len(o.indices)
            out["index"][i, :n] = o.indices + self.lo
            out["asr_sim"][i, :n] = o.asr_sim
            out["audio_sim"][i, :n] = o.audio_sim
            out["flags"][i, :n] = self.f[o.indices]
        return torch.from_numpy(out.view(np.uint8).reshape(len(queries), k, 24))

    def score_all(self, queries, class_weights):
        cls = no.legacy_good_speech(7, self.lo + len(self.f))[self.lo:]
        return np.stack([no.class_weight_scores(q, self.a, self.b, self.f & 1, self.f & 2, cls, class_weights)
                         for q in queries])

    def merge_candidates(self, gathered, w_asr, w_audio, k, threshold, to_host):
        w_asr, w_audio = self._w          # reuse the weights of the preceding search_candidates
        c = gathered.numpy().reshape(gathered.shape[0], gathered.shape[1], k * 24).view(CANDIDATE_DTYPE)
        res = []
        for i in range(c.shape[1]):
            rows = c[:, i, :].reshape(-1)
            rows = rows[rows["index"] >= 0]
            fusion, _, _ = no.fuse(rows["asr_sim"], rows["audio_sim"], rows["flags"], w_asr[i], w_audio[i])
            keep = fusion > threshold
            order = np.lexsort((rows["index"][keep], -fusion[keep]))[:k]
            res.append((rows["index"][keep][order], fusion[keep][order]))
        return res


def _worker(rank, world, port, seed, n_total, plants, nq, k, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_total, rank, world)
        idx = FakeShardIndex(seed, n_total, lo, hi, plants, nq)
        q = synth.raw_queries(seed, 0, nq)
        wa = np.array([0.5, 0.3, 0.68][:nq]); wb = 1.0 - wa
        out = ShardedSearcher(idx, rank, world).search(q, wa, wb, k=k, threshold=0.1)
        ret[rank] = [(i.tolist(), f.tolist()) for i, f in out]
        scores = ShardedSearcher(idx, rank, world).score_all(q, no.LEGACY_CLASS_WEIGHTS["adaptive"], n_total)
        ret[f"scores{rank}"] = scores.numpy()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_search_equals_single_library():
    import torch.multiprocessing as mp
    seed, n_total, plants, nq, k = 42, 3001, 25, 3, 20
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, seed, n_total, plants, nq, k, ret), nprocs=2, join=True)
    a, b, f, _ = synth.library(seed, n_total, nq, plants, True)
    q = synth.raw_queries(seed, 0, nq)
    wa = np.array([0.5, 0.3, 0.68]); wb = 1.0 - wa
    assert ret[0] == ret[1]                                   # every rank holds the merged result
    for i in range(nq):
        o = no.search(q[i], a, b, f, wa[i], wb[i], k=k)
        assert ret[0][i][0] == o.indices.tolist()
        np.testing.assert_allclose(ret[0][i][1], o.fusion, atol=1e-6, rtol=0)
    # legacy all-N scoring over the same two shards: every rank ends with the full vector
    cls = no.legacy_good_speech(7, n_total)
    assert ret["scores0"].shape == (nq, n_total) and np.array_equal(ret["scores0"], ret["scores1"])
    for i in range(nq):
        want = no.legacy_scores(q[i], a, b, f & 1, f & 2, cls, "adaptive")
        np.testing.assert_allclose(ret["scores0"][i], want, atol=1e-6, rtol=0)
