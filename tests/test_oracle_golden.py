"""The numpy oracle (oracle/numpy_oracle.py) against the golden fixtures minted from the
reference's own code (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from multimodal_audio_search_b200 import synth
from oracle import numpy_oracle as no

SCORE_TOL = 1e-6   # |delta score| <= 1e-6 oracle vs literal reference


def _check(res, rec, k=10):
    n = len(rec["indices"])
    assert list(res.indices[:n]) == rec["indices"]
    assert len(res.indices) == min(n, k) or n == k
    np.testing.assert_allclose(res.fusion[:n], rec["fusion"], atol=SCORE_TOL, rtol=0)
    np.testing.assert_allclose(res.asr_sim[:n], rec["asr_sim"], atol=SCORE_TOL, rtol=0)
    np.testing.assert_allclose(res.audio_sim[:n], rec["audio_sim"], atol=SCORE_TOL, rtol=0)
    np.testing.assert_array_equal(res.eff_asr_w[:n], rec["eff_asr_w"])
    np.testing.assert_array_equal(res.eff_audio_w[:n], rec["eff_audio_w"])


def test_search_cases(search_cases):
    for case in search_cases:
        a, b, f, _ = synth.library(case["seed"], case["n_rows"], case["n_queries"],
                                   case["plants"], case["partial"])
        q = synth.raw_queries(case["seed"], 0, case["n_queries"])
        k = case.get("k", 10)
        for rec in case["queries"]:
            res = no.search(q[rec["qi"]], a, b, f, rec["asr_weight"], rec["audio_weight"], k=k)
            _check(res, rec, k)
            if len(rec["indices"]) < k:      # the reference returned everything above 0.1
                assert len(res.indices) == len(rec["indices"])


def test_known_answer(known_answer):
    ka = known_answer
    for text, qkey in (("zzz", "q_zzz"), ("guitar solo", "q_guitar")):
        rec = ka["answers"][text]
        wa, wb, an = no.analyze_query_for_weights(text)
        assert (wa, wb, an) == (rec["asr_weight"], rec["audio_weight"], rec["analysis"])
        res = no.search(ka[qkey], ka["asr"], ka["audio"], ka["flags"], wa, wb,
                        has_asr=ka["has_asr"], has_audio=ka["has_audio"])
        _check(res, rec)
        assert len(res.indices) == len(rec["indices"])


def test_flag_cases(flag_cases):
    fc = flag_cases
    a, b, _, _ = synth.library(fc["seed"], fc["n_rows"], fc["n_queries"], fc["plants"], False)
    q = synth.raw_queries(fc["seed"], 0, fc["n_queries"])
    for rec in fc["queries"]:
        res = no.search(q[rec["qi"]], a, b, np.array(fc["flags"], np.uint8), rec["asr_weight"],
                        rec["audio_weight"], has_asr=np.array(fc["has_asr"], bool),
                        has_audio=np.array(fc["has_audio"], bool))
        _check(res, rec)
        assert len(res.indices) == len(rec["indices"])


def test_query_weights(query_weight_cases):
    for rec in query_weight_cases:
        got = no.analyze_query_for_weights(rec["query"])
        assert got == (rec["asr_weight"], rec["audio_weight"], rec["analysis"]), rec["query"]


def test_nonfinite_rejected():
    a, b, f, _ = synth.library(1, 8)
    a[3, 5] = np.nan
    with pytest.raises(ValueError):
        no.search(synth.raw_queries(1, 0, 1)[0], a, b, f, 0.5, 0.5)


def test_chunked_equals_whole():
    seed, n = 31, 5000
    a, b, f, _ = synth.library(seed, n, 1, 40, True)
    q = synth.raw_queries(seed, 0, 1)[0]
    whole = no.search(q, a, b, f, 0.3, 0.7, k=25)

    def src(r0, r1):
        x, y, g, _ = synth.library(seed, n, 1, 40, True, r0=r0, r1=r1)
        return x, y, g
    idx, fu, sa, sb = no.search_chunked(q, src, n, 0.3, 0.7, k=25, chunk=777)
    assert list(idx) == list(whole.indices)
    np.testing.assert_allclose(fu, whole.fusion, atol=1e-6, rtol=0)
