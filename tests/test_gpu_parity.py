"""GPU parity: the CUDA path (through the C-ABI, libcab.so) against the numpy oracle and the
golden fixtures minted from the reference.  Run on a B200 with `pytest -m gpu`."""
import numpy as np
import pytest

from multimodal_audio_search_b200 import SegmentIndex, synth, synth_queries
from oracle import numpy_oracle as no
from tests.util import BF16_TOL, FP32_TOL, assert_topk_matches, python_fusion, result_row

pytestmark = pytest.mark.gpu


def _index_from(a, b, f, dtype="fp32"):
    idx = SegmentIndex(dtype, capacity=len(f))
    idx.append(a, b, f)
    return idx


def _check_against_golden(res, rec, tol, q=0):
    gi, gf, ga, gb, gfl = result_row(res, q)
    assert list(gi) == rec["indices"]
    np.testing.assert_allclose(gf, rec["fusion"], atol=tol, rtol=0)
    np.testing.assert_allclose(ga, rec["asr_sim"], atol=tol, rtol=0)
    np.testing.assert_allclose(gb, rec["audio_sim"], atol=tol, rtol=0)
    for j in range(len(gi)):     # device float64 fusion == the reference's formula on its sims
        want, ea, eb = python_fusion(ga[j], gb[j], int(gfl[j]), rec["asr_weight"], rec["audio_weight"])
        assert gf[j] == want
        assert ea == rec["eff_asr_w"][j] and eb == rec["eff_audio_w"][j]


def test_device_synth_equals_numpy_synth():
    seed, n = 20261018, 3000
    a, b, f, _ = synth.library(seed, n, n_queries=4, plants=25, partial=True)
    idx = SegmentIndex("fp32")
    idx.append_synth(seed, n, 0, n, n_queries=4, plants=25, partial=True)
    assert len(idx) == n
    np.testing.assert_allclose(idx.read_rows(0, 0, n), no.normalize_rows(a), atol=2e-7, rtol=0)
    np.testing.assert_allclose(idx.read_rows(1, 0, n), no.normalize_rows(b), atol=2e-7, rtol=0)
    np.testing.assert_array_equal(synth_queries(seed, 0, 4), synth.raw_queries(seed, 0, 4))
    # flags travel through search results
    q = synth.raw_queries(seed, 0, 1)
    res = idx.search(q, 0.5, 0.5, k=25)
    gi, _, _, _, gfl = result_row(res)
    np.testing.assert_array_equal(gfl, f[gi])


@pytest.mark.parametrize("mode", ["ascending", "clustered"])
def test_device_synth_stress_distributions_equal_numpy(mode):
    """The two stress distributions bench.py measures: device generator == numpy twin, on a shard
    range in the middle of a 10 M-row global library (row numbers beyond 2^23)."""
    seed, n, r0, r1 = 20261018, 10_000_000, 9_990_000, 9_992_000
    a, b, f, _ = synth.library(seed, n, partial=True, r0=r0, r1=r1, mode=mode)
    idx = SegmentIndex("fp32")
    idx.append_synth(seed, n, r0, r1, partial=True, mode=mode)
    np.testing.assert_allclose(idx.read_rows(0, 0, r1 - r0), no.normalize_rows(a), atol=2e-7, rtol=0)
    np.testing.assert_allclose(idx.read_rows(1, 0, r1 - r0), no.normalize_rows(b), atol=2e-7, rtol=0)
    np.testing.assert_array_equal(idx.read_flags(), f)
    q = synth.bench_queries(seed, mode, 0, 2)
    res = idx.search(q, 0.5, 0.5, k=10)
    for qi in range(2):
        o = no.search(q[qi], a, b, f, 0.5, 0.5, k=10)
        gi, gf, _, _, _ = result_row(res, qi)
        assert_topk_matches(gi, gf, o, FP32_TOL)
    idx.close()


@pytest.mark.parametrize("dtype,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_golden_search_cases(search_cases, dtype, tol):
    for case in search_cases:
        a, b, f, _ = synth.library(case["seed"], case["n_rows"], case["n_queries"], case["plants"], case["partial"])
        q = synth.raw_queries(case["seed"], 0, case["n_queries"])
        idx = _index_from(a, b, f, dtype)
        k = case.get("k", 10)
        for rec in case["queries"]:
            res = idx.search(q[rec["qi"]], rec["asr_weight"], rec["audio_weight"], k=k)
            if dtype == "fp32":
                _check_against_golden(res, rec, tol)
            else:
                o = no.search(q[rec["qi"]], a, b, f, rec["asr_weight"], rec["audio_weight"], k=k)
                gi, gf, _, _, _ = result_row(res)
                assert_topk_matches(gi, gf, o, tol)
        # all queries of the case in one call (per-query weights)
        recs = case["queries"]
        res = idx.search(q[[r["qi"] for r in recs]], [r["asr_weight"] for r in recs],
                         [r["audio_weight"] for r in recs], k=k)
        if dtype == "fp32":
            for i, rec in enumerate(recs):
                _check_against_golden(res, rec, tol, q=i)


def test_known_answer(known_answer):
    ka = known_answer
    asr = np.where(ka["has_asr"][:, None], ka["asr"], 0).astype(np.float32)      # None -> zero row
    aud = np.where(ka["has_audio"][:, None], ka["audio"], 0).astype(np.float32)
    idx = _index_from(asr, aud, ka["flags"])
    for text, qkey in (("zzz", "q_zzz"), ("guitar solo", "q_guitar")):
        rec = ka["answers"][text]
        res = idx.search(ka[qkey], rec["asr_weight"], rec["audio_weight"])
        gi, gf, ga, gb, _ = result_row(res)
        # seg5/seg7 sit within 1e-8 of the 0.1 threshold in the reference: straddlers allowed
        o = no.search(ka[qkey], ka["asr"], ka["audio"], ka["flags"], rec["asr_weight"], rec["audio_weight"],
                      has_asr=ka["has_asr"], has_audio=ka["has_audio"])
        assert_topk_matches(gi, gf, o, FP32_TOL)
        assert list(gi[:6]) == rec["indices"][:6]          # incl. the exact tie seg0 < seg1


def test_flag_cases(flag_cases):
    fc = flag_cases
    a, b, _, _ = synth.library(fc["seed"], fc["n_rows"], fc["n_queries"], fc["plants"], False)
    q = synth.raw_queries(fc["seed"], 0, fc["n_queries"])
    a = np.where(np.array(fc["has_asr"], bool)[:, None], a, 0).astype(np.float32)
    b = np.where(np.array(fc["has_audio"], bool)[:, None], b, 0).astype(np.float32)
    idx = _index_from(a, b, np.array(fc["flags"], np.uint8))
    for rec in fc["queries"]:
        res = idx.search(q[rec["qi"]], rec["asr_weight"], rec["audio_weight"])
        _check_against_golden(res, rec, FP32_TOL)


def test_edge_cases():
    idx = SegmentIndex("fp32")
    q = synth.raw_queries(1, 0, 1)
    res = idx.search(q)                                   # empty index
    assert res.count[0] == 0 and (res.indices == -1).all()
    a, b, f, _ = synth.library(1, 5, 1, 3)
    idx.append(a, b, f)
    res = idx.search(q, k=128)                            # k larger than the library
    o = no.search(q[0], a, b, f, 0.5, 0.5, k=128)
    assert list(result_row(res)[0]) == list(o.indices)
    with pytest.raises(ValueError):                       # sklearn: "Input contains NaN"
        bad = q.copy(); bad[0, 7] = np.nan
        idx.search(bad)
    res = idx.search(q)                                   # handle still usable
    assert list(result_row(res)[0]) == list(o.indices[:10])
    with pytest.raises(ValueError):
        bad = a.copy(); bad[2, 0] = np.inf
        idx.append(bad, b, f)
    assert len(idx) == 5                                  # nothing appended
    zq = np.zeros((1, 384), np.float32)                   # zero query: all cosines 0 -> no results
    assert idx.search(zq).count[0] == 0


@pytest.mark.parametrize("n", [1, 31, 257, 4099])
def test_ragged_sizes_and_incremental_append(n):
    seed = 50 + n
    a, b, f, _ = synth.library(seed, n, 1, min(n, 20), True)
    q = synth.raw_queries(seed, 0, 1)
    o = no.search(q[0], a, b, f, 0.35, 0.65, k=16)
    for dtype, tol in (("fp32", FP32_TOL), ("bf16", BF16_TOL)):
        idx = SegmentIndex(dtype)                          # capacity 0: grows as it goes
        cuts = sorted({0, n // 3, (2 * n) // 3, n})
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            idx.append(a[lo:hi], b[lo:hi], f[lo:hi])
        assert len(idx) == n
        res = idx.search(q, 0.35, 0.65, k=16)
        gi, gf, _, _, _ = result_row(res)
        assert_topk_matches(gi, gf, o, tol)


def test_scan_variants_agree():
    seed, n = 77, 20000
    idx = SegmentIndex("fp32")
    idx.append_synth(seed, n, 0, n, n_queries=2, plants=40, partial=True)
    q = synth.raw_queries(seed, 0, 2)
    base = None
    idx.set_option("gemv_query_tile", 1)              # one query per pass: exercises the (U, MB) variants
    for unroll, bps in ((8, 1), (4, 2), (4, 1), (2, 3), (2, 2), (1, 4), (1, 2)):
        idx.set_option("gemv_unroll", unroll)
        idx.set_option("gemv_blocks_per_sm", bps)
        res = idx.search(q, [0.5, 0.3], [0.5, 0.7], k=50)
        key = (res.indices.tolist(), res.fusion.tolist(), res.count.tolist())
        base = base or key
        assert key == base


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_query_tiles_agree(dtype):
    """Scoring 1, 2 or 4 queries per corpus pass (register tiling) gives bit-identical results,
    also for batches that do not fill the last tile."""
    seed, n, nq = 78, 50000, 7
    idx = SegmentIndex(dtype)
    idx.append_synth(seed, n, 0, n, n_queries=nq, plants=40, partial=True)
    q = synth.raw_queries(seed, 0, nq)
    wa = np.linspace(0.2, 0.8, nq); wb = 1.0 - wa
    base = None
    for tile in (1, 2, 4, 0):
        idx.set_option("gemv_query_tile", tile)
        for k, thr in ((10, 0.1), (100, -1.0)):
            res = idx.search(q, wa, wb, k=k, threshold=thr, path="gemv")
            key = (res.indices.tolist(), res.fusion.tolist(), res.asr_sim.tolist(), res.count.tolist())
            base = base or {}
            assert base.setdefault((k, thr), key) == key, (dtype, tile, k)


def test_finalize_paths_agree():
    """Head-bound fast path vs the general streaming/bitonic path of the finalize kernel."""
    seed, n = 91, 200000
    for dtype in ("fp32", "bf16"):
        idx = SegmentIndex(dtype)
        idx.append_synth(seed, n, 0, n, n_queries=3, plants=150, partial=True)
        q = synth.raw_queries(seed, 0, 3)
        for k in (1, 10, 37, 100, 128):
            for thr in (0.1, -1.0):                 # -1: every row passes -> full candidate lists
                idx.set_option("finalize_general", 0)
                fast = idx.search(q, [0.5, 0.2, 0.8], [0.5, 0.8, 0.2], k=k, threshold=thr)
                idx.set_option("finalize_general", 1)
                gen = idx.search(q, [0.5, 0.2, 0.8], [0.5, 0.8, 0.2], k=k, threshold=thr)
                assert fast.indices.tolist() == gen.indices.tolist(), (dtype, k, thr)
                assert fast.fusion.tolist() == gen.fusion.tolist()
                assert fast.count.tolist() == gen.count.tolist()
                if thr < 0:
                    assert (fast.count == k).all()


def test_threshold_minus_one_matches_oracle():
    """With the threshold out of the way every row competes: exercises full partial lists."""
    seed, n, k = 17, 30000, 64
    a, b, f, _ = synth.library(seed, n, 1, 0, True)
    q = synth.raw_queries(seed, 0, 1)
    o = no.search(q[0], a, b, f, 0.4, 0.6, k=k, threshold=-1.0)
    idx = _index_from(a, b, f)
    res = idx.search(q, 0.4, 0.6, k=k, threshold=-1.0)
    gi, gf, _, _, _ = result_row(res)
    assert_topk_matches(gi, gf, o, FP32_TOL, threshold=-1.0)


def test_top100_peel_case(search_cases):
    case = [c for c in search_cases if c.get("k") == 100][0]
    a, b, f, _ = synth.library(case["seed"], case["n_rows"], 1, case["plants"], case["partial"])
    q = synth.raw_queries(case["seed"], 0, 1)
    idx = _index_from(a, b, f)
    res = idx.search(q, 0.5, 0.5, k=100)
    _check_against_golden(res, case["queries"][0], FP32_TOL)


def test_device_tensor_interface():
    torch = pytest.importorskip("torch")
    seed, n = 12, 5000
    idx = SegmentIndex("fp32")
    idx.append_synth(seed, n, 0, n, n_queries=3, plants=30)
    q = synth.raw_queries(seed, 0, 3)
    host = idx.search(q, [0.5, 0.2, 0.8], [0.5, 0.8, 0.2], k=20)
    dev = idx.search(torch.from_numpy(q).cuda(), [0.5, 0.2, 0.8], [0.5, 0.8, 0.2], k=20)
    # no device-wide synchronize: .cpu() must be ordered after the search on torch's current stream
    np.testing.assert_array_equal(dev.indices.cpu().numpy(), host.indices)
    np.testing.assert_array_equal(dev.fusion.cpu().numpy(), host.fusion)
    np.testing.assert_array_equal(dev.count.cpu().numpy(), host.count)
    # device-side append of raw rows == host append
    a, b, f, _ = synth.library(seed, n, 3, 30)
    idx2 = SegmentIndex("fp32")
    idx2.append(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(f).cuda())
    host2 = idx2.search(q, [0.5, 0.2, 0.8], [0.5, 0.8, 0.2], k=20)
    np.testing.assert_array_equal(host2.indices, host.indices)


def test_device_calls_are_ordered_on_the_callers_stream():
    """Results of CUDA-tensor calls must be valid for stream-ordered consumers (no device sync),
    on torch's default stream and on a side stream, for a scan long enough to expose a race."""
    torch = pytest.importorskip("torch")
    seed, n, nq = 8, 2_000_000, 512
    idx = SegmentIndex("bf16", capacity=n)
    idx.append_synth(seed, n, 0, n, n_queries=nq, plants=20)
    q = synth.raw_queries(seed, 0, nq)
    wa = np.full(nq, 0.4); wb = 1.0 - wa
    host = idx.search(q, wa, wb, k=10, path="gemm")
    qd = torch.from_numpy(q).cuda()
    dev = idx.search(qd, wa, wb, k=10, path="gemm")
    np.testing.assert_array_equal(dev.count.cpu().numpy(), host.count)
    np.testing.assert_array_equal(dev.indices.cpu().numpy(), host.indices)
    host_gemv = idx.search(q[:40], wa[:40], wb[:40], k=10, path="gemv")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        dev2 = idx.search(qd[:40], wa[:40], wb[:40], k=10, path="gemv")
        got = dev2.indices.cpu().numpy()
    np.testing.assert_array_equal(got, host_gemv.indices)


@pytest.mark.parametrize("k", [10, 100])
def test_sharded_merge_equals_single_index(k):
    torch = pytest.importorskip("torch")
    seed, n = 321, 30000
    whole = SegmentIndex("fp32")
    whole.append_synth(seed, n, 0, n, n_queries=2, plants=80, partial=True)
    q = synth.raw_queries(seed, 0, 2)
    wa, wb = [0.5, 0.68], [0.5, 0.32]
    ref = whole.search(q, wa, wb, k=k)
    cuts = [0, 7001, 15000, 22222, n]
    blocks = []
    shards = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        s = SegmentIndex("fp32")
        s.append_synth(seed, n, lo, hi, n_queries=2, plants=80, partial=True)
        s.row_base = lo
        shards.append(s)
        blocks.append(s.search_candidates(q, wa, wb, k=k))
    gathered = torch.stack(blocks).contiguous()
    out = shards[0].merge_candidates(gathered, wa, wb, k=k)
    np.testing.assert_array_equal(out.indices, ref.indices)
    np.testing.assert_array_equal(out.fusion, ref.fusion)
    np.testing.assert_array_equal(out.count, ref.count)


def test_config2_one_million_fp32():
    """BASELINE.json configs[1]: 1M-segment dual corpus, single query, fp32, top-10."""
    seed, n = 20261018, 1_000_000
    idx = SegmentIndex("fp32", capacity=n)
    idx.append_synth(seed, n, 0, n, n_queries=1, plants=30)
    q = synth.raw_queries(20261018, 0, 1)

    def src(r0, r1):
        a, b, f, _ = synth.library(seed, n, 1, 30, False, r0=r0, r1=r1)
        return a, b, f
    oi, of, osa, osb = no.search_chunked(q[0], src, n, 0.5, 0.5, k=10, chunk=50000)
    res = idx.search(q, 0.5, 0.5, k=10)
    gi, gf, ga, gb, _ = result_row(res)
    assert list(gi) == list(oi)                       # planted neighbours: gaps >> 1e-5
    np.testing.assert_allclose(gf, of, atol=FP32_TOL, rtol=0)
    np.testing.assert_allclose(ga, osa, atol=FP32_TOL, rtol=0)
    np.testing.assert_allclose(gb, osb, atol=FP32_TOL, rtol=0)
    # size-independent properties: idempotence, sortedness, threshold
    res2 = idx.search(q, 0.5, 0.5, k=10)
    assert res2.indices.tolist() == res.indices.tolist() and res2.fusion.tolist() == res.fusion.tolist()
    assert all(gf[i] >= gf[i + 1] for i in range(len(gf) - 1)) and gf.min() > 0.1
    # bf16 storage of the same library: recall@10 and tolerance
    idb = SegmentIndex("bf16", capacity=n)
    idb.append_synth(seed, n, 0, n, n_queries=1, plants=30)
    rb = idb.search(q, 0.5, 0.5, k=10)
    bi, bf, _, _, _ = result_row(rb)
    assert len(set(bi) & set(oi)) >= 10
    np.testing.assert_allclose(np.sort(bf)[::-1], of, atol=BF16_TOL, rtol=0)


@pytest.mark.parametrize("dtype,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_zero_query_weight_skips_rows_that_only_have_that_pipeline(dtype, tol):
    """Weights (1, 0) / (0, 1) -- single-corpus search, as in the earlier engines.  A row whose only
    successful pipeline carries weight 0 has effective weights summing to 0 and is skipped
    (audio_search.py:659-661); it must not crowd real results out of the scan's top-k either."""
    seed, n = 77, 20000
    a, b, f, _ = synth.library(seed, n, 2, 30, True)
    q = synth.raw_queries(seed, 0, 2)
    # make the single-pipeline rows the best matches of the corpus that will carry weight 0
    only_audio = np.nonzero(f == 2)[0][:200]
    only_asr = np.nonzero(f == 1)[0][:200]
    b[only_audio] = q[0] + 0.05 * b[only_audio]
    a[only_asr] = q[0] + 0.05 * a[only_asr]
    idx = _index_from(a, b, f, dtype)
    for wa, wb in ((1.0, 0.0), (0.0, 1.0), (1e-300, 1.0)):
        for k in (10, 100):
            res = idx.search(q, wa, wb, k=k)
            for qi in range(2):
                o = no.search(q[qi], a, b, f, wa, wb, k=k)
                gi, gf, _, _, gfl = result_row(res, qi)
                assert_topk_matches(gi, gf, o, tol, k=k)
                if wb == 0.0:
                    assert not np.isin(gi, only_audio).any() and (gfl & 1).all()
                if wa == 0.0:
                    assert not np.isin(gi, only_asr).any() and (gfl & 2).all()
    with pytest.raises(Exception, match="must be > 0"):
        idx.search(q, 0.0, 0.0)
    idx.close()


def test_invalid_arguments_are_rejected_not_computed():
    import ctypes as C
    from multimodal_audio_search_b200 import _native as N
    idx = SegmentIndex("fp32")
    a, b, f, _ = synth.library(2, 40, 1, 5)
    idx.append(a, b, f)
    q = synth.raw_queries(2, 0, 1)
    for kwargs in ({"k": 0}, {"k": 129}, {"threshold": float("nan")}, {"path": "gemm"}):
        with pytest.raises((N.CabError, KeyError)):
            idx.search(q, 0.5, 0.5, **kwargs)
    with pytest.raises(N.CabError):
        idx.search(q, -0.1, 0.5)                         # weights must be >= 0
    with pytest.raises(ValueError):
        idx.search(np.zeros((1, 100), np.float32))       # wrong dimension
    with pytest.raises(ValueError):
        idx.append(a[:, :100], b, f)
    big = np.zeros((4097, 384), np.float32)
    with pytest.raises(N.CabError):
        idx.search(big)                                  # > CAB_MAX_QUERIES per call
    lib = N.lib()
    assert lib.cab_search(idx._h, None, 0, None, None, 1, 10, 0.1, 0, None, None, None, None, None, None, 0, None) == N.CAB_ERR_INVALID
    assert b"null" in lib.cab_last_error(idx._h)
    res = idx.search(q)                                  # the handle is still healthy
    assert res.count[0] >= 1
    with pytest.raises(N.CabError):
        SegmentIndex("fp32", device=99)
