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


# ---- host logic of the two drop-ins with an oracle-backed fake index (no GPU) ----------------------
class _FakeIndex:
    def __init__(self, dtype="fp32", capacity=0, device=0):
        self.a = np.zeros((0, 384), np.float32); self.b = np.zeros((0, 384), np.float32); self.f = np.zeros(0, np.uint8)
        self.appends = 0

    def __len__(self):
        return len(self.f)

    def append(self, asr, audio, flags):
        audio = np.zeros_like(asr) if audio is None else audio
        self.a, self.b, self.f = np.vstack([self.a, asr]), np.vstack([self.b, audio]), np.concatenate([self.f, flags])
        self.appends += 1

    def clear(self):
        self.a, self.b, self.f = self.a[:0], self.b[:0], self.f[:0]

    def set_option(self, key, value):
        self.options = getattr(self, "options", {}) | {key: value}      # unit-length rows here: raw dot == cosine

    def pinned_scores(self, nq=1):
        return np.empty((nq, len(self)), np.float32)

    def score_all(self, queries, class_weights, out=None):
        res = np.stack([no.class_weight_scores(q, self.a, self.b, self.f & 1, self.f & 2, self.f >> 2, class_weights)
                        for q in np.atleast_2d(queries)])
        if out is not None:
            out[...] = res
            return out
        return res

    def search(self, queries, w_asr, w_audio, k=10, threshold=0.1, path="auto"):
        from multimodal_audio_search_b200.index import SearchResult
        q = np.atleast_2d(queries)
        o = no.search(q[0], self.a, self.b, self.f, w_asr, w_audio, k=k, threshold=threshold)
        idx = np.full((1, k), -1, np.int64); idx[0, :len(o.indices)] = o.indices
        z = np.zeros((1, k), np.float32)
        return SearchResult(idx, np.zeros((1, k)), z, z, np.zeros((1, k), np.uint8), np.array([len(o.indices)], np.int32))


def test_drop_in_host_logic_with_a_fake_index(monkeypatch):
    monkeypatch.setattr(legacy, "SegmentIndex", _FakeIndex)
    z, meta = _cases()
    m = meta[0]
    a, b, f, _ = synth.library(m["seed"], m["n_rows"], m["n_queries"], m["plants"], m["partial"])
    q = synth.raw_queries(m["seed"], 0, m["n_queries"])
    good = no.legacy_good_speech(m["seed"], m["n_rows"])
    db = rs.legacy_database(a, b, (f & 1).astype(bool), (f & 2).astype(bool), good)
    eng = legacy.UnifiedAudioSearch(sentence_model=rs.ListEmbedder({"q0": q[0], "q1": q[1]}))
    assert eng.search("q0", [], "adaptive").shape == (0,)
    grown = list(db[:200])
    assert eng.search("q0", grown, "asr_only").shape == (200,)
    grown.extend(db[200:])
    for qi in range(2):
        for strategy in ("asr_only", "caption_only", "adaptive"):
            sims = eng.search(f"q{qi}", grown, strategy)
            assert sims.dtype == np.float64 and np.abs(sims - z[f"{m['name']}/{qi}/{strategy}"]).max() <= TOL
    assert eng._cab_database.index.appends == 2                       # incremental sync, no rebuild
    eng.search("q0", list(db), "adaptive")                            # another list object: rebuilt
    assert eng._cab_database.index.appends == 3 and len(eng._cab_database.index) == len(db)

    # clean_audio_search.py drop-in
    case = json.load(open(CLEAN_GOLD))[0]
    a, c, mm, ha, hc, q, qm = no.clean_library(case["seed"], case["n_rows"], 2, case["plants"])
    cdb = rs.clean_database(a, c, mm, ha, hc)
    texts = {f"{r['mode']} {r['qi']}": (qm if r["mode"] == "combined" else q)[r["qi"]] for r in case["queries"]}
    ceng = legacy.CleanAudioSearch(text_embedder=rs.FakeEmbedder(texts))
    assert ceng.search_audio("asr 0", "asr") == []
    ceng.audio_database.extend(cdb)
    for r in case["queries"]:
        got = ceng.search_audio(f"{r['mode']} {r['qi']}", r["mode"])
        assert [int(x["segment_id"][4:]) for x in got] == r["indices"]
        np.testing.assert_allclose([x["similarity"] for x in got], r["similarity"], atol=1e-6, rtol=0)
    # embeddings of any length are accepted: the engine is told to rank by the raw dot product (:306)
    seg = dict(cdb[0]); seg["asr_embedding"] = 3.0 * cdb[0]["combined_embedding"]
    ceng.audio_database.append(seg)
    ceng.search_audio("asr 0", "asr")
    assert ceng._cab_database.pair.options == {"raw_dot": 1} and ceng._cab_database.combined.options == {"raw_dot": 1}
    assert len(ceng._cab_database.pair) == len(cdb) + 1
