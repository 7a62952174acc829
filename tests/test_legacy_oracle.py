"""CPU: the oracle's restatement of the earlier engine's modes
(previous_iterations/streamlit_app.py:173-223) against the golden vectors minted from the
reference (tests/golden/legacy_scores.npz, oracle/make_golden.py) and, in the build container,
against the reference itself; plus the host logic of the drop-in (flag / class bytes)."""
import json
import os

import numpy as np
import pytest

from multimodal_audio_search_b200 import legacy, synth
from oracle import numpy_oracle as no
from oracle import reference_shim as rs

GOLD = os.path.join(os.path.dirname(__file__), "golden", "legacy_scores.npz")
TOL = 1e-5          # fp32 similarities, BASELINE.json north_star tolerance


def _cases():
    z = np.load(GOLD)
    return z, json.loads(str(z["meta"]))


def test_oracle_matches_golden_vectors():
    z, meta = _cases()
    assert len(meta) >= 2
    for m in meta:
        a, b, f, _ = synth.library(m["seed"], m["n_rows"], m["n_queries"], m["plants"], m["partial"])
        q = synth.raw_queries(m["seed"], 0, m["n_queries"])
        good = no.legacy_good_speech(m["seed"], m["n_rows"])
        assert 0 < good.sum() < m["n_rows"]
        for qi in range(m["n_queries"]):
            for strategy in ("asr_only", "caption_only", "adaptive"):
                want = z[f"{m['name']}/{qi}/{strategy}"]
                got = no.legacy_scores(q[qi], a, b, f & 1, f & 2, good, strategy)
                assert want.shape == (m["n_rows"],) and got.dtype == np.float32
                assert np.abs(got.astype(np.float64) - want).max() <= TOL
                # a missing embedding scores exactly 0.0 in its single-corpus mode
                if strategy == "asr_only":
                    assert np.all(want[(f & 1) == 0] == 0.0)
                # planted neighbours rank first under the reference's own argsort (:410)
                if strategy == "adaptive" and m["plants"]:
                    top_ref = np.argsort(want)[::-1][:3]
                    top_got = np.argsort(got)[::-1][:3]
                    assert list(top_ref) == list(top_got)


def test_class_weight_form_equals_the_strategies():
    rng = np.random.default_rng(4)
    a = rng.standard_normal((200, 384)).astype(np.float32)
    b = rng.standard_normal((200, 384)).astype(np.float32)
    q = rng.standard_normal(384).astype(np.float32)
    has_a, has_b = rng.random(200) > 0.2, rng.random(200) > 0.2
    cls = rng.integers(0, 2, 200)
    for strategy in ("asr_only", "caption_only", "adaptive"):
        want = no.legacy_scores(q, a, b, has_a, has_b, cls, strategy)
        got = no.class_weight_scores(q, a, b, has_a, has_b, cls, no.LEGACY_CLASS_WEIGHTS[strategy])
        assert np.array_equal(want, got)
        assert legacy.CLASS_WEIGHTS[strategy] == no.LEGACY_CLASS_WEIGHTS[strategy]


def test_item_flags():
    e = np.ones(384, np.float32)
    assert legacy.item_flags({"asr_embedding": e, "caption_embedding": e, "asr_transcription": "hello there world"}) == 3 | 4
    assert legacy.item_flags({"asr_embedding": e, "caption_embedding": None, "asr_transcription": "  short     "}) == 1
    assert legacy.item_flags({"asr_embedding": None, "caption_embedding": e}) == 2
    assert legacy.item_flags({"asr_embedding": None, "caption_embedding": None, "asr_transcription": "x" * 11}) == 4
    assert legacy.item_flags({"asr_embedding": e, "caption_embedding": e, "asr_transcription": "x" * 10}) == 3


@pytest.mark.skipif(not rs.legacy_available(), reason="reference not mounted (GPU box)")
def test_oracle_matches_the_reference_live():
    rng = np.random.default_rng(9)
    n = 150
    a = rng.standard_normal((n, 384)).astype(np.float32)
    b = rng.standard_normal((n, 384)).astype(np.float32)
    q = rng.standard_normal(384).astype(np.float32)
    a[:10] = q + 0.4 * a[:10]
    b[5:15] = 3.0 * (q + 0.6 * b[5:15])
    has_a, has_b, good = rng.random(n) > 0.25, rng.random(n) > 0.25, rng.random(n) > 0.5
    db = rs.legacy_database(a, b, has_a, has_b, good)
    assert [legacy.item_flags(it) for it in db] == \
        [int(has_a[i]) | int(has_b[i]) << 1 | int(good[i]) << 2 for i in range(n)]
    for strategy in ("asr_only", "caption_only", "adaptive", "anything else"):
        ref = rs.legacy_search("a query", q, db, strategy)
        got = no.legacy_scores(q, a, b, has_a, has_b, good, strategy if strategy in no.LEGACY_CLASS_WEIGHTS else "adaptive")
        assert ref.shape == (n,)
        assert np.abs(ref - got.astype(np.float64)).max() <= TOL
    assert rs.legacy_search("a query", q, [], "adaptive").shape == (0,)


# ---- previous_iterations/clean_audio_search.py: search_audio (:293-320) ----------------------------
CLEAN_GOLD = os.path.join(os.path.dirname(__file__), "golden", "clean_search.json")
_MODE_ROWS = {"asr": 0, "caption": 1, "combined": 2}


def test_clean_search_oracle_matches_golden_vectors():
    cases = json.load(open(CLEAN_GOLD))
    assert len(cases) >= 2
    for case in cases:
        a, c, m, ha, hc, q, qm = no.clean_library(case["seed"], case["n_rows"], 2, case["plants"])
        everything = np.ones(case["n_rows"], bool)
        for rec in case["queries"]:
            if rec["mode"] not in _MODE_ROWS:
                assert rec["indices"] == []                       # unknown mode: every similarity 0.0
                continue
            rows, has, qv = {"asr": (a, ha, q), "caption": (c, hc, q), "combined": (m, everything, qm)}[rec["mode"]]
            idx, sims = no.clean_search(qv[rec["qi"]], rows, has)
            assert idx.tolist() == rec["indices"]
            np.testing.assert_allclose(sims, rec["similarity"], atol=1e-6, rtol=0)
            assert len(idx) <= 10 and (sims > 0.1).all() and all(sims[j] >= sims[j + 1] for j in range(len(sims) - 1))


@pytest.mark.skipif(not rs.clean_available(), reason="reference not mounted (GPU box)")
def test_clean_search_oracle_matches_the_reference_live():
    a, c, m, ha, hc, q, qm = no.clean_library(91, 300, 2, 9)
    db = rs.clean_database(a, c, m, ha, hc)
    for mode, rows, has, qv in (("asr", a, ha, q[0]), ("caption", c, hc, q[1]), ("combined", m, np.ones(300, bool), qm[0])):
        ref = rs.clean_search("text", qv, db, mode)
        idx, sims = no.clean_search(qv, rows, has)
        assert [int(r["segment_id"][4:]) for r in ref] == idx.tolist()
        assert [r["similarity"] for r in ref] == sims.tolist()     # same expression, same floats
        assert all(set(r) == set(db[0]) | {"similarity"} for r in ref)
    assert rs.clean_search("text", q[0], [], "asr") == []
