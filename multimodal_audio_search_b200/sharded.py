"""Corpus sharded by segment over the GPUs of one box (one process per GPU, torch.distributed).

Each rank owns the contiguous global segment range `shard_range(n_total, rank, world)` in its own
`SegmentIndex` (with `row_base` = first global segment), so a search is:

  1. local fused scan -> this shard's top-k as packed 24-byte candidates   (libcab, on device)
  2. ONE exchange of the candidate blocks  [world, Q, k, 24]  (<= 24 B x k x Q per rank:
     latency-bound, the path's only exchange step), either
       exchange="p2p"  (default on GPUs): fused into the finalize kernel -- every rank stores its
                       candidates straight into every rank's buffer over NVLink peer memory
                       (CUDA IPC) and raises an epoch flag; the merge kernel waits on the flags;
       exchange="nccl": torch.distributed.all_gather_into_tensor (also what the gloo CPU test uses)
  3. merge on every rank: float64 reference fusion, threshold, (score desc, global index asc),
     first k                                                              (libcab, on device)

Queries and weights are replicated (1.5 KB per query).  The reference has no counterpart: it is a
single-process loop over one Python list (audio_search.py:639).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

CANDIDATE_DTYPE = np.dtype([("index", "<i8"), ("asr_sim", "<f4"), ("audio_sim", "<f4"),
                            ("flags", "<u4"), ("pad", "<u4")])      # cab_candidate, 24 bytes


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, near-equal split: rank r owns [r*ceil(n/world), ...) clipped to n_total."""
    per = -(-n_total // world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


class ShardedSearcher:
    """Drives steps 1-3 for one rank.  `index` is this rank's SegmentIndex (or any object with
    `search_candidates` / `merge_candidates`); `group` a torch.distributed process group."""

    def __init__(self, index, rank: int = 0, world: int = 1, group=None, exchange: str = "nccl",
                 max_queries: int = 256, max_k: int = 128):
        self.index, self.rank, self.world, self.group = index, rank, world, group
        self._gathered = None
        self.exchange = exchange if world > 1 else "nccl"
        if self.exchange not in ("nccl", "p2p"):
            raise ValueError("exchange must be 'nccl' or 'p2p'")
        self.max_queries, self.max_k = max_queries, max_k
        if self.exchange == "p2p":
            import torch.distributed as dist
            handle = index.peer_init(rank, world, max_queries, max_k)
            handles = [None] * world
            dist.all_gather_object(handles, handle, group=group)      # host-side, once
            index.peer_attach(b"".join(handles))
            dist.barrier(group=group)

    def search(self, queries, w_asr, w_audio, k: int = 10, threshold: float = 0.1, path: str = "auto",
               to_host: bool = True):
        import torch
        import torch.distributed as dist
        if self.exchange == "p2p":
            n = queries.shape[0] if hasattr(queries, "shape") and len(queries.shape) == 2 else 1
            if n <= self.max_queries and k <= self.max_k:
                return self.index.search_sharded(queries, w_asr, w_audio, k=k, threshold=threshold, path=path,
                                                 to_host=to_host)
        cands = self.index.search_candidates(queries, w_asr, w_audio, k=k, threshold=threshold, path=path)
        if self.world == 1:
            gathered = cands.unsqueeze(0)
        else:
            # concatenated layout [world*Q, k, 24] (accepted by both NCCL and gloo), viewed per rank
            shape = (self.world * cands.shape[0],) + tuple(cands.shape[1:])
            if self._gathered is None or tuple(self._gathered.shape) != shape or self._gathered.device != cands.device:
                self._gathered = torch.empty(shape, dtype=cands.dtype, device=cands.device)
            dist.all_gather_into_tensor(self._gathered, cands.contiguous(), group=self.group)
            gathered = self._gathered.view((self.world,) + tuple(cands.shape))
        # the weights are already on the device (staged by search_candidates on the same handle)
        return self.index.merge_candidates(gathered, None, None, k=k, threshold=threshold, to_host=to_host)

    def score_all(self, queries, class_weights, n_total: int):
        """Legacy all-N scoring (cab_score_all) over a sharded library: every rank scores its own
        rows, one all-gather of the score vectors (4 bytes per segment and query) gives every rank
        the full [Q, n_total] result.  Shards are `shard_range(n_total, rank, world)`; the blocks
        are padded to the largest shard for the collective.  Returns a tensor on the index's
        device (numpy if the index returns numpy and world == 1)."""
        import torch
        import torch.distributed as dist
        local = self.index.score_all(queries, class_weights)
        if self.world == 1:
            return local
        if not torch.is_tensor(local):
            local = torch.from_numpy(np.ascontiguousarray(local))
        nq = local.shape[0]
        per = -(-n_total // self.world)
        lo, hi = shard_range(n_total, self.rank, self.world)
        if local.shape[1] != hi - lo:
            raise ValueError(f"rank {self.rank} holds {local.shape[1]} rows, its shard of {n_total} is {hi - lo}")
        block = torch.zeros((nq, per), dtype=local.dtype, device=local.device)
        block[:, :hi - lo] = local
        gathered = torch.empty((self.world * nq, per), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(gathered, block, group=self.group)
        # [world, Q, per] -> [Q, world*per]: rank r's block starts at its first global row r*per, so
        # position == global segment index and only the tail beyond n_total is padding
        full = gathered.view(self.world, nq, per).permute(1, 0, 2).reshape(nq, self.world * per)
        return full[:, :n_total].contiguous()
