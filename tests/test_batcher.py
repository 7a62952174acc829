"""CPU: the micro-batching front end (multimodal_audio_search_b200/batcher.py) with an oracle-backed
fake index -- coalescing, grouping by (k, threshold), per-request results, error isolation,
exclusive calls, shutdown.  The GPU run of the same logic is tests/test_gpu_engine.py."""
import threading
import time

import numpy as np
import pytest

from multimodal_audio_search_b200 import SearchBatcher, synth
from multimodal_audio_search_b200.index import SearchResult
from oracle import numpy_oracle as no


class FakeIndex:
    """`.search` computed by the numpy oracle (test infrastructure); records the batch sizes."""

    def __init__(self, seed=5, n=400, delay=0.0):
        self.a, self.b, self.f, _ = synth.library(seed, n, 4, 8, True)
        self.calls, self.delay, self.thread_ids = [], delay, set()

    def search(self, queries, w_asr, w_audio, k=10, threshold=0.1, path="auto"):
        self.thread_ids.add(threading.get_ident())
        time.sleep(self.delay)
        q = np.atleast_2d(queries)
        wa = np.broadcast_to(np.asarray(w_asr, dtype=np.float64), (len(q),))
        wb = np.broadcast_to(np.asarray(w_audio, dtype=np.float64), (len(q),))
        self.calls.append((len(q), k, threshold))
        if not np.isfinite(q).all():
            raise ValueError("Input contains NaN or infinity (query)")
        idx = np.full((len(q), k), -1, np.int64)
        fus = np.zeros((len(q), k))
        cnt = np.zeros(len(q), np.int32)
        for i in range(len(q)):
            o = no.search(q[i], self.a, self.b, self.f, wa[i], wb[i], k=k, threshold=threshold)
            c = len(o.indices)
            idx[i, :c], fus[i, :c], cnt[i] = o.indices, o.fusion, c
        z = np.zeros((len(q), k), np.float32)
        return SearchResult(idx, fus, z, z.copy(), np.zeros((len(q), k), np.uint8), cnt)


def test_concurrent_requests_share_gpu_calls_and_get_their_own_rows():
    fake = FakeIndex(delay=0.02)
    q = synth.raw_queries(5, 0, 4)
    weights = [(0.5, 0.5), (0.3, 0.7), (0.8, 0.2), (0.44, 0.56)]
    results = {}
    with SearchBatcher(fake, max_batch=64, max_wait_s=0.05) as batcher:
        def session(t):
            qi = t % 4
            results[t] = batcher.search(q[qi], *weights[qi], k=10 if t % 2 == 0 else 5)
        threads = [threading.Thread(target=session, args=(t,)) for t in range(24)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        stats = batcher.stats
    assert stats.requests == 24 and stats.batches < 24 and stats.largest_batch > 1
    assert sum(c[0] for c in fake.calls) == 24 and {c[1] for c in fake.calls} == {5, 10}    # grouped by k
    assert len(fake.thread_ids) == 1                                                         # one caller of the handle
    for t, res in results.items():
        qi, k = t % 4, 10 if t % 2 == 0 else 5
        o = no.search(q[qi], fake.a, fake.b, fake.f, *weights[qi], k=k)
        assert res.indices.shape == (1, k) and int(res.count[0]) == len(o.indices)
        assert res.indices[0, :len(o.indices)].tolist() == o.indices.tolist()
        assert res.fusion[0, :len(o.indices)].tolist() == o.fusion.tolist()


def test_a_bad_query_fails_alone():
    fake = FakeIndex(delay=0.01)
    q = synth.raw_queries(5, 0, 3)
    bad = q[1].copy(); bad[7] = np.nan
    with SearchBatcher(fake, max_wait_s=0.05) as batcher:
        futs = [batcher.submit(q[0], 0.5, 0.5), batcher.submit(bad, 0.5, 0.5), batcher.submit(q[2], 0.5, 0.5)]
        assert int(futs[0].result().count[0]) >= 0 and int(futs[2].result().count[0]) >= 0
        with pytest.raises(ValueError, match="NaN"):
            futs[1].result()
        with pytest.raises(ValueError):
            batcher.submit(np.zeros(100, np.float32), 0.5, 0.5)           # wrong length: rejected at submit


def test_exclusive_calls_run_on_the_worker_between_batches_and_close_is_clean():
    fake = FakeIndex()
    q = synth.raw_queries(5, 0, 1)
    batcher = SearchBatcher(fake, max_batch=2)
    seen = batcher.run_exclusive(lambda: threading.get_ident())
    batcher.search(q[0], 0.5, 0.5)
    assert seen in fake.thread_ids and seen != threading.get_ident()
    with pytest.raises(ZeroDivisionError):
        batcher.run_exclusive(lambda: 1 / 0)
    assert int(batcher.search(q[0], 0.5, 0.5).count[0]) >= 0               # still alive after a failing call
    futs = [batcher.submit(q[0], 0.5, 0.5) for _ in range(5)]              # max_batch=2: 3 GPU calls
    batcher.close()
    assert all(f.done() for f in futs) and max(c[0] for c in fake.calls) <= 2
    with pytest.raises(RuntimeError):
        batcher.submit(q[0], 0.5, 0.5)
    with pytest.raises(ValueError):
        SearchBatcher(fake, max_batch=0)
