"""GPU: the tcgen05 / TMEM batched path (path="gemm", bf16 storage) against the GEMV path on the
same index and against the fp32 numpy oracle (tolerance 2e-3, recall@10 >= 0.999)."""
import numpy as np
import pytest

from multimodal_audio_search_b200 import SegmentIndex, synth
from oracle import numpy_oracle as no
from tests.util import BF16_TOL, result_row

pytestmark = pytest.mark.gpu

W_CLASSES = [0.5, 0.2, 0.3, 0.4, 0.6, 0.7, 0.8]


def _weights(nq):
    wa = np.array([W_CLASSES[i % len(W_CLASSES)] for i in range(nq)])
    return wa, 1.0 - wa


def _compare_paths(idx, q, wa, wb, k, threshold=0.1, slack=1e-3):
    """gemm selects with bf16-rounded queries (score error ~1e-4) and both paths re-score winners
    identically, so results may differ only among candidates within `slack` of the k-th score."""
    gm = idx.search(q, wa, wb, k=k, threshold=threshold, path="gemm")
    wide = idx.search(q, wa, wb, k=min(128, k + 28), threshold=threshold, path="gemv")
    for i in range(q.shape[0]):
        gi, gf, ga, gb, _ = result_row(gm, i)
        vi, vf, va, vb, _ = result_row(wide, i)
        score_of = dict(zip(vi.tolist(), vf.tolist()))
        n = min(len(vi), k)
        assert len(gi) == n or abs(len(gi) - n) <= 2, (i, len(gi), n)
        kth = vf[n - 1] if n else threshold
        for j, (row, f) in enumerate(zip(gi.tolist(), gf.tolist())):
            assert row in score_of, (i, j, row, f, kth)
            assert f == score_of[row]                      # identical re-scoring
            if j < n and row != vi[j]:
                assert abs(f - vf[j]) <= slack, (i, j, row, vi[j], f, vf[j])
        missing = set(vi[:n].tolist()) - set(gi.tolist())
        for row in missing:
            assert abs(score_of[row] - kth) <= slack, (i, row, score_of[row], kth)
        assert all(gf[j] >= gf[j + 1] for j in range(len(gf) - 1))


@pytest.mark.parametrize("n,nq,k", [(128, 64, 10), (5000, 64, 10), (40000, 256, 100), (33333, 100, 37),
                                    (20000, 300, 10), (64, 8, 5)])
def test_gemm_matches_gemv(n, nq, k):
    seed = 1000 + n
    idx = SegmentIndex("bf16")
    idx.append_synth(seed, n, 0, n, n_queries=min(nq, max(1, n // 50)), plants=min(40, n // 4), partial=True)
    q = synth.raw_queries(seed, 0, nq)
    wa, wb = _weights(nq)
    _compare_paths(idx, q, wa, wb, k)


def test_gemm_every_row_competes():
    """threshold -1: every row passes, candidate lists fill and are compacted repeatedly."""
    seed, n, nq, k = 77, 30000, 128, 100
    idx = SegmentIndex("bf16")
    idx.append_synth(seed, n, 0, n, n_queries=4, plants=50, partial=True)
    q = synth.raw_queries(seed, 0, nq)
    wa, wb = _weights(nq)
    _compare_paths(idx, q, wa, wb, k, threshold=-1.0)
    res = idx.search(q, wa, wb, k=k, threshold=-1.0, path="gemm")
    assert (res.count == k).all()


def test_gemm_vs_fp32_oracle_recall():
    seed, n, nq, k = 20261018, 60000, 64, 10
    a, b, f, _ = synth.library(seed, n, nq, 30, False)
    q = synth.raw_queries(seed, 0, nq)
    idx = SegmentIndex("bf16", capacity=n)
    idx.append(a, b, f)
    wa, wb = _weights(nq)
    res = idx.search(q, wa, wb, k=k, path="gemm")
    hit = tot = 0
    for i in range(nq):
        o = no.search(q[i], a, b, f, wa[i], wb[i], k=k)
        gi, gf, _, _, _ = result_row(res, i)
        hit += len(set(gi.tolist()) & set(o.indices.tolist()))
        tot += len(o.indices)
        common = min(len(gi), len(o.indices))
        np.testing.assert_allclose(gf[:common], o.fusion[:common], atol=BF16_TOL, rtol=0)
    assert hit / tot >= 0.999, hit / tot


def test_auto_path_uses_gemm_for_batches():
    seed, n = 5, 10000
    idx = SegmentIndex("bf16")
    idx.append_synth(seed, n, 0, n, n_queries=8, plants=30)
    q = synth.raw_queries(seed, 0, 80)
    wa, wb = _weights(80)
    auto = idx.search(q, wa, wb, k=10)                      # >= 64 queries on bf16 -> tensor cores
    gm = idx.search(q, wa, wb, k=10, path="gemm")
    assert auto.indices.tolist() == gm.indices.tolist()
    f32 = SegmentIndex("fp32")
    f32.append_synth(seed, n, 0, n, n_queries=8, plants=30)
    with pytest.raises(Exception):
        f32.search(q, wa, wb, k=10, path="gemm")            # tensor-core path needs bf16 storage
