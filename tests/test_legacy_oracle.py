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
