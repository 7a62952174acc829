"""Micro-batching front end: many sessions, one library, one GPU call per batch.

The reference answers one query per `search_with_fusion` call (audio_search.py:624), one engine
per browser session (:708-711).  When several sessions (threads) search the SAME library, their
queries can share a corpus pass: 4 queries per pass in the GEMV scan, and from 4 queries on a bf16 (or shadowed fp32)
index the tensor-core scan ("makes Q >= 64 batches arise naturally
from concurrent sessions").  `SearchBatcher` collects the requests that arrive while the GPU is
busy (or within `max_wait_s` of the first one) and issues ONE `SegmentIndex.search` for them.

A `SegmentIndex` handle is externally synchronised (Threading); the batcher's
worker thread is the only caller of the index it owns, so any number of threads may call
`search()` / `submit()`.
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .index import SearchResult

DIM = 384
_NOTHING = object()


@dataclass
class _Request:
    query: np.ndarray
    w_asr: float
    w_audio: float
    k: int
    threshold: float
    future: Future


@dataclass
class _Call:
    fn: object
    future: Future


@dataclass
class BatcherStats:
    requests: int = 0
    batches: int = 0
    largest_batch: int = 0

    @property
    def mean_batch(self) -> float:
        return self.requests / self.batches if self.batches else 0.0


class SearchBatcher:
    """`index`: a SegmentIndex (anything with `.search(queries, w_asr, w_audio, k=, threshold=,
    path=)`).  `max_batch` <= 4096 (CAB_MAX_QUERIES); `max_wait_s`: how long the first request of
    a batch may wait for company (0 = only coalesce what queued up while the GPU was busy)."""

    def __init__(self, index, max_batch: int = 256, max_wait_s: float = 0.0, path: str = "auto"):
        if not 1 <= max_batch <= 4096:
            raise ValueError("max_batch must be in 1..4096")
        self.index, self.max_batch, self.max_wait_s, self.path = index, int(max_batch), float(max_wait_s), path
        self.stats = BatcherStats()
        self._queue: "queue.Queue[Optional[_Request]]" = queue.Queue()
        self._closed = False
        self._pushed_back = _NOTHING               # item taken off the queue while collecting, handled next
        self._worker = threading.Thread(target=self._run, name="cab-search-batcher", daemon=True)
        self._worker.start()

    # ---- client side (any thread) ------------------------------------------------------------
    def submit(self, query, w_asr: float, w_audio: float, k: int = 10, threshold: float = 0.1) -> Future:
        """Queue one query; the Future resolves to a SearchResult with one row."""
        if self._closed:
            raise RuntimeError("the batcher is closed")
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        if q.shape[0] != DIM:
            raise ValueError(f"expected a {DIM}-element query, got {q.shape[0]}")
        fut: Future = Future()
        self._queue.put(_Request(q, float(w_asr), float(w_audio), int(k), float(threshold), fut))
        return fut

    def search(self, query, w_asr: float, w_audio: float, k: int = 10, threshold: float = 0.1) -> SearchResult:
        """Blocking form of `submit` (what a session thread calls)."""
        return self.submit(query, w_asr, w_audio, k, threshold).result()

    def run_exclusive(self, fn):
        """Run `fn()` on the worker thread between two batches and return its result: the way to
        touch the index (append segments, save, change options) while sessions are searching."""
        if self._closed:
            raise RuntimeError("the batcher is closed")
        fut: Future = Future()
        self._queue.put(_Call(fn, fut))
        return fut.result()

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._queue.put(None)
            self._worker.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- worker ------------------------------------------------------------------------------
    def _collect(self, first: _Request) -> List[_Request]:
        batch = [first]
        deadline = time.monotonic() + self.max_wait_s
        while len(batch) < self.max_batch:
            try:
                timeout = deadline - time.monotonic()
                item = self._queue.get_nowait() if timeout <= 0 else self._queue.get(timeout=timeout)
            except queue.Empty:
                break
            if item is None or isinstance(item, _Call):    # close() / exclusive call: finish this batch first
                self._pushed_back = item
                break
            batch.append(item)
        return batch

    def _run(self) -> None:
        while True:
            first, self._pushed_back = (self._pushed_back, _NOTHING) if self._pushed_back is not _NOTHING \
                else (self._queue.get(), _NOTHING)
            if first is None:
                return
            if isinstance(first, _Call):
                try:
                    first.future.set_result(first.fn())
                except BaseException as e:
                    first.future.set_exception(e)
                continue
            batch = self._collect(first)
            self.stats.requests += len(batch)
            # one GPU call per (k, threshold) group, request order kept inside a group
            groups = {}
            for r in batch:
                groups.setdefault((r.k, r.threshold), []).append(r)
            for (k, threshold), reqs in groups.items():
                self.stats.batches += 1
                self.stats.largest_batch = max(self.stats.largest_batch, len(reqs))
                try:
                    res = self.index.search(np.stack([r.query for r in reqs]),
                                            np.array([r.w_asr for r in reqs]), np.array([r.w_audio for r in reqs]),
                                            k=k, threshold=threshold, path=self.path)
                except BaseException as e:         # a bad query (NaN) fails its whole group: retry one by one
                    if len(reqs) == 1:
                        reqs[0].future.set_exception(e)
                    else:
                        for r in reqs:
                            try:
                                r.future.set_result(self.index.search(r.query[None, :], r.w_asr, r.w_audio, k=k,
                                                                      threshold=threshold, path=self.path))
                            except BaseException as e1:
                                r.future.set_exception(e1)
                    continue
                for i, r in enumerate(reqs):
                    r.future.set_result(SearchResult(res.indices[i:i + 1], res.fusion[i:i + 1], res.asr_sim[i:i + 1],
                                                     res.audio_sim[i:i + 1], res.flags[i:i + 1], res.count[i:i + 1]))
