"""GPU: BASELINE.json's full sizes (10 M-segment dual corpus) checked through size-independent
properties, without scanning 30 GB on the CPU:

* planted-neighbour ground truth: the oracle scores ONLY the planted rows of a query (a few dozen
  rows regenerated on the host); every planted row that beats the engine's k-th score must be in
  the engine's result, at the oracle's score;
* sortedness, threshold, idempotence, (score desc, index asc) order;
* tensor-core path == GEMV path on the same bf16 index (up to the k-th boundary)."""
import numpy as np
import pytest

from multimodal_audio_search_b200 import SegmentIndex, synth
from oracle import numpy_oracle as no
from tests.util import BF16_TOL, FP32_TOL, result_row

pytestmark = pytest.mark.gpu

N_ROWS, SEED, PLANTS = 10_000_000, 20261018, 24


def _planted_truth(qi, q, wa, wb, n_queries):
    """Oracle fused scores of query qi's planted rows: {global row: (fusion, s_asr, s_audio)}."""
    spec = synth.plant_spec(SEED, N_ROWS, n_queries, PLANTS)
    out = {}
    for r in synth.plant_rows_of_query(spec, qi).tolist():
        a, b, f, _ = synth.library(SEED, N_ROWS, n_queries, PLANTS, False, r0=r, r1=r + 1)
        o = no.search(q, a, b, f, wa, wb, k=1, threshold=-1.0)
        out[r] = (float(o.all_fusion[0]), float(no.cosine_rows(q, a)[0]), float(no.cosine_rows(q, b)[0]))
    return out


def _check_properties(res, i, threshold=0.1):
    gi, gf, _, _, _ = result_row(res, i)
    assert len(set(gi.tolist())) == len(gi) and (gf > threshold).all()
    assert all(gf[j] > gf[j + 1] or (gf[j] == gf[j + 1] and gi[j] < gi[j + 1]) for j in range(len(gf) - 1))
    return gi, gf


def test_10m_fp32_single_query_top10():
    idx = SegmentIndex("fp32", capacity=N_ROWS)
    idx.append_synth(SEED, N_ROWS, 0, N_ROWS, n_queries=4, plants=PLANTS)
    q = synth.raw_queries(SEED, 0, 4)
    for qi, (wa, wb) in enumerate(((0.5, 0.5), (0.30000000000000004, 0.7), (0.8, 0.19999999999999996), (0.2, 0.8))):
        res = idx.search(q[qi], wa, wb, k=10)
        again = idx.search(q[qi], wa, wb, k=10)
        assert res.indices.tolist() == again.indices.tolist() and res.fusion.tolist() == again.fusion.tolist()
        gi, gf = _check_properties(res, 0)
        assert len(gi) == 10
        truth = _planted_truth(qi, q[qi], wa, wb, 4)
        got = dict(zip(gi.tolist(), zip(gf.tolist(), res.asr_sim[0].tolist(), res.audio_sim[0].tolist())))
        kth = gf[-1]
        for row, (fu, sa, sb) in truth.items():
            if fu > kth + FP32_TOL:
                assert row in got, (qi, row, fu, kth)
            if row in got:
                assert abs(got[row][0] - fu) <= FP32_TOL and abs(got[row][1] - sa) <= FP32_TOL and abs(got[row][2] - sb) <= FP32_TOL
        assert len(set(gi.tolist()) & set(truth)) >= 8              # the planted rows dominate the top-10
    idx.close()


def test_10m_bf16_batch256_top100_tensor_cores():
    nq, k = 256, 100
    idx = SegmentIndex("bf16", capacity=N_ROWS)
    idx.append_synth(SEED, N_ROWS, 0, N_ROWS, n_queries=nq, plants=PLANTS)
    q = synth.raw_queries(SEED, 0, nq)
    wa = np.array([[0.5, 0.2, 0.3, 0.4, 0.6, 0.7, 0.8][i % 7] for i in range(nq)]); wb = 1.0 - wa
    gm = idx.search(q, wa, wb, k=k, path="gemm")
    again = idx.search(q, wa, wb, k=k, path="gemm")
    assert gm.indices.tolist() == again.indices.tolist()
    assert (gm.count == k).all()
    for i in (0, 1, 100, 255):
        gi, gf = _check_properties(gm, i)
        truth = _planted_truth(i, q[i], wa[i], wb[i], nq)
        got = dict(zip(gi.tolist(), gf.tolist()))
        kth = gf[-1]
        for row, (fu, _, _) in truth.items():
            if fu > kth + BF16_TOL:
                assert row in got, (i, row, fu, kth)
            if row in got:
                assert abs(got[row] - fu) <= BF16_TOL
    gv = idx.search(q[:32], wa[:32], wb[:32], k=k, path="gemv")
    hits = sum(len(set(gm.indices[i].tolist()) & set(gv.indices[i].tolist())) for i in range(32))
    assert hits / (32 * k) >= 0.99                                   # differences only at the k-th boundary
    idx.close()


def test_10m_legacy_scores_are_bit_exact_combinations():
    """cab_score_all at BASELINE size (10 M segments, bf16 rows): size-independent properties.
    asr_only / caption_only return the two cosines themselves, so every weighted mode must be
    BIT-identical to numpy's float32 `w_a * s_asr + w_b * s_caption` of those two vectors with the
    row's class weights (previous_iterations/streamlit_app.py:217/:219); planted rows carry the
    oracle's score; the top-k search agrees with the score vector where both rules coincide."""
    n_queries = 2
    idx = SegmentIndex("bf16", capacity=N_ROWS)
    idx.append_synth(SEED, N_ROWS, 0, N_ROWS, n_queries=n_queries, plants=PLANTS, partial=False)
    cls = no.legacy_good_speech(SEED, N_ROWS).astype(np.uint8) * 1 + (np.arange(N_ROWS) % 7 == 0).astype(np.uint8) * 2
    idx.set_weight_classes(cls)
    flags = idx.read_flags()
    assert np.array_equal(flags >> 2, cls) and np.all((flags & 3) == 3)
    q = synth.raw_queries(SEED, 0, n_queries)
    s_asr = idx.score_all(q, no.LEGACY_CLASS_WEIGHTS["asr_only"])
    s_cap = idx.score_all(q, no.LEGACY_CLASS_WEIGHTS["caption_only"])
    assert s_asr.shape == (n_queries, N_ROWS) and np.isfinite(s_asr).all() and np.abs(s_asr).max() <= 1.0 + 1e-3
    table = np.array([[0.2, 0.8], [0.7, 0.3], [-0.5, 1.5], [0.25, 0.0]], dtype=np.float32)
    got = idx.score_all(q, table)
    for qi in range(n_queries):
        want = table[cls, 0] * s_asr[qi] + table[cls, 1] * s_cap[qi]          # float32, op by op
        assert want.dtype == np.float32
        assert np.array_equal(got[qi], want)
        truth = _planted_truth(qi, q[qi], 0.5, 0.5, n_queries)                  # oracle cosines of the planted rows
        for r, (_, sa, sb) in truth.items():
            assert abs(s_asr[qi, r] - sa) <= BF16_TOL and abs(s_cap[qi, r] - sb) <= BF16_TOL
        # idempotence
        assert np.array_equal(idx.score_all(q[qi], table)[0], got[qi])
    # both pipelines successful + weights summing to 1: the current engine's fusion is the same
    # linear form, so its top-10 must be the 10 largest entries of the score vector
    half = idx.score_all(q, [[0.5, 0.5]] * 4)
    res = idx.search(q, 0.5, 0.5, k=10, threshold=0.1, path="gemv")
    for qi in range(n_queries):
        gi, gf = _check_properties(res, qi)
        top = np.argpartition(-half[qi], 10)[:10]
        kth = np.sort(half[qi][top])[0]
        assert np.abs(half[qi][gi] - gf).max() <= 1e-6
        assert all(half[qi][i] >= kth - 1e-6 for i in gi)
    idx.close()
