"""GPU: score distributions that defeat threshold / bound pruning (SURVEY.md section 8(d) secondary
distributions): monotonically ASCENDING scores (every segment beats all earlier ones, so every
running bound is useless) and CLUSTERED corpora (thousands of segments above the 0.1 threshold,
like anisotropic sentence-embedding spaces).  fp32 GEMV vs the oracle; bf16 GEMM vs bf16 GEMV."""
import numpy as np
import pytest

from multimodal_audio_search_b200 import SegmentIndex
from oracle import numpy_oracle as no
from tests.util import FP32_TOL, assert_topk_matches, result_row

pytestmark = pytest.mark.gpu


def _unit(x):
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def _ascending(n, nq, rng):
    q = _unit(rng.standard_normal((nq, 384))).astype(np.float32)
    t = np.linspace(0.15, 0.97, n)[:, None]                  # cosine to query 0 grows with the row index
    noise = _unit(rng.standard_normal((n, 384)))
    a = (t * q[0] + np.sqrt(1 - t * t) * noise).astype(np.float32)
    b = (t * q[0] + np.sqrt(1 - t * t) * _unit(rng.standard_normal((n, 384)))).astype(np.float32)
    return q, a, b, np.full(n, 3, np.uint8)


def _clustered(n, nq, rng, n_centres=32, spread=0.5):
    centres = _unit(rng.standard_normal((n_centres, 384)))
    which = rng.integers(0, n_centres, n)
    a = _unit(centres[which] + spread * _unit(rng.standard_normal((n, 384)))).astype(np.float32)
    b = _unit(centres[which] + spread * _unit(rng.standard_normal((n, 384)))).astype(np.float32)
    q = _unit(centres[rng.integers(0, n_centres, nq)] + 0.3 * _unit(rng.standard_normal((nq, 384)))).astype(np.float32)
    f = rng.choice(np.array([1, 2, 3, 3, 3, 3], np.uint8), n)
    a[(f & 1) == 0] = 0
    b[(f & 2) == 0] = 0
    return q, a, b, f


@pytest.mark.parametrize("maker,k", [(_ascending, 10), (_ascending, 100), (_clustered, 10), (_clustered, 128)])
def test_gemv_fp32_vs_oracle(maker, k):
    rng = np.random.default_rng(7)
    q, a, b, f = maker(30000, 3, rng)
    idx = SegmentIndex("fp32", capacity=len(f))
    idx.append(a, b, f)
    wa, wb = np.array([0.5, 0.3, 0.8]), np.array([0.5, 0.7, 0.2])
    res = idx.search(q, wa, wb, k=k)
    for i in range(3):
        o = no.search(q[i], a, b, f, wa[i], wb[i], k=k)
        gi, gf, _, _, _ = result_row(res, i)
        assert (o.all_fusion > 0.1).sum() > 5 * k or i > 0          # the distribution really floods the threshold
        assert_topk_matches(gi, gf, o, FP32_TOL)


@pytest.mark.parametrize("maker,k", [(_ascending, 10), (_ascending, 100), (_clustered, 100)])
def test_gemm_bf16_vs_gemv_bf16(maker, k):
    rng = np.random.default_rng(11)
    nq = 96
    q, a, b, f = maker(40000, nq, rng)
    idx = SegmentIndex("bf16", capacity=len(f))
    idx.append(a, b, f)
    wa = np.linspace(0.2, 0.8, nq); wb = 1 - wa
    gm = idx.search(q, wa, wb, k=k, path="gemm")
    wide = idx.search(q, wa, wb, k=min(128, k + 28), path="gemv")
    for i in range(nq):
        gi, gf, _, _, _ = result_row(gm, i)
        vi, vf, _, _, _ = result_row(wide, i)
        score_of = dict(zip(vi.tolist(), vf.tolist()))
        n = min(len(vi), k)
        assert abs(len(gi) - n) <= 2
        kth = vf[n - 1] if n else 0.1
        for row, fu in zip(gi.tolist(), gf.tolist()):
            assert row in score_of and fu == score_of[row]
        for row in set(vi[:n].tolist()) - set(gi.tolist()):
            assert abs(score_of[row] - kth) <= 1e-3
