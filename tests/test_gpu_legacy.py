"""GPU: `cab_score_all` (all-N fused similarities with per-row weight classes) and the drop-in for
the earlier engine's `UnifiedAudioSearch.search` (previous_iterations/streamlit_app.py:173-223)
against the golden vectors minted from the reference and the numpy oracle."""
import json
import os

import numpy as np
import pytest

from multimodal_audio_search_b200 import SegmentIndex, legacy, synth
from oracle import numpy_oracle as no
from oracle.reference_shim import ListEmbedder, legacy_database

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "legacy_scores.npz")
TOL = {"fp32": 1e-5, "bf16": 2e-3}      # BASELINE.json north_star tolerances


def _library(m):
    a, b, f, _ = synth.library(m["seed"], m["n_rows"], m["n_queries"], m["plants"], m["partial"])
    q = synth.raw_queries(m["seed"], 0, m["n_queries"])
    good = no.legacy_good_speech(m["seed"], m["n_rows"])
    return a, b, f, q, good


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_score_all_matches_reference_vectors(dtype):
    z = np.load(GOLD)
    for m in json.loads(str(z["meta"])):
        a, b, f, q, good = _library(m)
        idx = SegmentIndex(dtype, device=0)
        idx.append(a, b, (f | (good.astype(np.uint8) << 2)).astype(np.uint8))
        for strategy in ("asr_only", "caption_only", "adaptive"):
            got = idx.score_all(q, no.LEGACY_CLASS_WEIGHTS[strategy])
            assert got.shape == (m["n_queries"], m["n_rows"]) and got.dtype == np.float32
            for qi in range(m["n_queries"]):
                want = z[f"{m['name']}/{qi}/{strategy}"]
                assert np.abs(got[qi].astype(np.float64) - want).max() <= TOL[dtype]
                if strategy == "asr_only":
                    assert np.all(got[qi][(f & 1) == 0] == 0.0)          # missing embedding: exactly 0.0
        # the class bits are invisible to the top-k search (same result as without them)
        plain = SegmentIndex(dtype, device=0)
        plain.append(a, b, f)
        r0, r1 = plain.search(q, 0.6, 0.4, k=10), idx.search(q, 0.6, 0.4, k=10)
        assert np.array_equal(r0.indices, r1.indices) and np.array_equal(r0.fusion, r1.fusion)
        assert np.array_equal(r1.flags[r1.indices >= 0] & 3, r0.flags[r0.indices >= 0])
        idx.close(); plain.close()


@pytest.mark.parametrize("n", [1, 3, 7, 31, 33, 64, 257, 1000, 4099])
def test_ragged_sizes_and_general_class_weights(n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, 384)).astype(np.float32)
    b = rng.standard_normal((n, 384)).astype(np.float32) * 5.0
    has_a, has_b = rng.random(n) > 0.2, rng.random(n) > 0.2
    a[~has_a] = 0; b[~has_b] = 0
    cls = rng.integers(0, 4, n)
    flags = (has_a.astype(np.uint8) | (has_b.astype(np.uint8) << 1) | (cls.astype(np.uint8) << 2))
    q = rng.standard_normal((3, 384)).astype(np.float32)
    cw = rng.uniform(-1, 1, (3, 4, 2)).astype(np.float32)
    for dtype in ("fp32", "bf16"):
        idx = SegmentIndex(dtype, device=0)
        idx.append(a, b, flags)
        got = idx.score_all(q, cw)
        for qi in range(3):
            want = no.class_weight_scores(q[qi], a, b, has_a, has_b, cls, cw[qi])
            assert np.abs(got[qi] - want).max() <= TOL[dtype]
        idx.close()


def test_device_tensors_nan_query_and_empty_index():
    import torch
    rng = np.random.default_rng(0)
    n = 5000
    a = rng.standard_normal((n, 384)).astype(np.float32)
    b = rng.standard_normal((n, 384)).astype(np.float32)
    q = rng.standard_normal((2, 384)).astype(np.float32)
    idx = SegmentIndex("fp32", device=0)
    assert idx.score_all(q, no.LEGACY_CLASS_WEIGHTS["adaptive"]).shape == (2, 0)
    idx.append(a, b, None)
    host = idx.score_all(q, no.LEGACY_CLASS_WEIGHTS["adaptive"])
    dev = idx.score_all(torch.from_numpy(q).cuda(), no.LEGACY_CLASS_WEIGHTS["adaptive"])
    torch.cuda.synchronize()
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), host)         # same kernel, same bits
    bad = q.copy(); bad[1, 5] = np.nan
    with pytest.raises(ValueError, match="NaN or infinity"):
        idx.score_all(bad, no.LEGACY_CLASS_WEIGHTS["adaptive"])
    again = idx.score_all(q, no.LEGACY_CLASS_WEIGHTS["adaptive"])           # the handle stays usable
    assert np.array_equal(again, host)
    dev_bad = idx.score_all(torch.from_numpy(bad).cuda(), no.LEGACY_CLASS_WEIGHTS["adaptive"])
    torch.cuda.synchronize()
    assert torch.isnan(dev_bad[1]).all() and np.array_equal(dev_bad[0].cpu().numpy(), host[0])
    res = idx.search(q, 0.5, 0.5, k=5)                                      # and so does the search
    assert res.count.shape == (2,)
    with pytest.raises(ValueError):
        idx.score_all(q, np.zeros((3, 2), np.float32))
    pinned = idx.pinned_scores(2)                                            # page-locked output buffer
    assert idx.score_all(q, no.LEGACY_CLASS_WEIGHTS["adaptive"], out=pinned) is pinned
    assert np.array_equal(pinned, host)
    with pytest.raises(ValueError, match="out must be"):
        idx.score_all(q, no.LEGACY_CLASS_WEIGHTS["adaptive"], out=np.zeros((2, n), np.float64))
    idx.close()


def test_scores_agree_with_the_topk_search():
    """Size-independent property at a larger size: with both pipelines successful and weights
    summing to 1 the current engine's fusion (audio_search.py:654-670) equals the class-weight
    form, so the top-k of the score vector must be the search's top-k."""
    n, seed = 300_000, 31
    idx = SegmentIndex("fp32", capacity=n, device=0)
    idx.append_synth(seed, n, 0, n, n_queries=2, plants=16, partial=False)
    q = synth.raw_queries(seed, 0, 2)
    scores = idx.score_all(q, [[0.3, 0.7]] * 4)
    res = idx.search(q, 0.3, 0.7, k=10, threshold=0.1)
    for qi in range(2):
        c = int(res.count[qi])
        assert c > 0
        order = np.lexsort((np.arange(n), -scores[qi].astype(np.float64)))[:c]
        assert np.abs(scores[qi][res.indices[qi, :c]] - res.fusion[qi, :c]).max() <= 1e-6
        assert set(order) == set(res.indices[qi, :c]) or \
            np.abs(np.sort(scores[qi][order]) - np.sort(scores[qi][res.indices[qi, :c]])).max() <= 1e-6
    idx.close()


def test_drop_in_for_the_earlier_engine():
    z = np.load(GOLD)
    m = json.loads(str(z["meta"]))[0]
    a, b, f, q, good = _library(m)
    db = legacy_database(a, b, (f & 1).astype(bool), (f & 2).astype(bool), good)
    texts = {f"query {qi}": q[qi] for qi in range(m["n_queries"])}
    eng = legacy.UnifiedAudioSearch(sentence_model=ListEmbedder(texts))
    assert eng.search("query 0", [], "adaptive").shape == (0,)
    half = len(db) // 2
    grown = list(db[:half])
    first = eng.search("query 0", grown, "adaptive")
    assert first.shape == (half,) and first.dtype == np.float64
    grown.extend(db[half:])                                                  # :338 appends
    for qi in range(m["n_queries"]):
        for strategy in ("asr_only", "caption_only", "adaptive"):
            sims = eng.search(f"query {qi}", grown, strategy)
            want = z[f"{m['name']}/{qi}/{strategy}"]
            assert sims.dtype == np.float64 and sims.shape == want.shape
            assert np.abs(sims - want).max() <= 1e-5
            assert list(eng.top_indices(sims, 3)) == list(np.argsort(want)[::-1][:3])
    assert np.array_equal(eng.search("query 1", grown, "no such strategy"), eng.search("query 1", grown, "adaptive"))

    class ReferenceLike:
        def __init__(self):
            self.sentence_model = ListEmbedder(texts)

        def search(self, *a, **k):
            raise AssertionError("CPU path must not run")
    patched = legacy.accelerate_legacy(ReferenceLike())
    sims = patched.search("query 1", db, "caption_only")
    assert np.abs(sims - z[f"{m['name']}/1/caption_only"]).max() <= 1e-5


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_drop_in_for_clean_audio_search(dtype):
    """`search_audio(query, search_mode)` of previous_iterations/clean_audio_search.py:293-320 on
    the GPU: same segment ids, and -- re-scored on the host with the reference's expression -- the
    same similarity floats as the oracle computes here (golden values within 1e-6)."""
    from oracle.reference_shim import FakeEmbedder, clean_database
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "clean_search.json")))
    for case in cases:
        a, c, m, ha, hc, q, qm = no.clean_library(case["seed"], case["n_rows"], 2, case["plants"])
        db = clean_database(a, c, m, ha, hc)
        texts = {}
        for rec in case["queries"]:
            texts[f"{rec['mode']} {rec['qi']}"] = (qm if rec["mode"] == "combined" else q)[rec["qi"]]
        eng = legacy.CleanAudioSearch(dtype=dtype, text_embedder=FakeEmbedder(texts))
        assert eng.search_audio("asr 0", "asr") == []                             # empty database (:295)
        half = len(db) // 2
        eng.audio_database.extend(db[:half])
        eng.search_audio("asr 0", "asr")
        eng.audio_database.extend(db[half:])                                      # grows like the app's list
        everything = np.ones(case["n_rows"], bool)
        for rec in case["queries"]:
            got = eng.search_audio(f"{rec['mode']} {rec['qi']}", rec["mode"])
            assert [int(r["segment_id"][4:]) for r in got] == rec["indices"]
            np.testing.assert_allclose([r["similarity"] for r in got], rec["similarity"], atol=1e-6, rtol=0)
            if rec["mode"] in ("asr", "caption", "combined"):
                rows, has, qv = {"asr": (a, ha, q), "caption": (c, hc, q), "combined": (m, everything, qm)}[rec["mode"]]
                _, sims = no.clean_search(qv[rec["qi"]], rows, has)
                assert [r["similarity"] for r in got] == sims.tolist()            # the reference's own floats
            for r in got:
                assert set(r) == set(db[0]) | {"similarity"} and type(r["similarity"]) is float
    # non-unit embeddings (and a non-unit query): :306 ranks by the RAW dot product, which is not the
    # cosine ranking -- served through the index's row lengths, same ids and floats as the oracle
    rng = np.random.default_rng(3)
    scale_a = rng.uniform(0.2, 5.0, len(db)).astype(np.float32)
    scale_c = rng.uniform(0.2, 5.0, len(db)).astype(np.float32)
    a2, c2, m2 = a * scale_a[:, None], c * scale_c[:, None], m * scale_a[:, None]
    db2 = clean_database(a2, c2, m2, ha, hc)
    q2 = (1.7 * q[0]).astype(np.float32)
    raw = legacy.CleanAudioSearch(dtype=dtype, text_embedder=FakeEmbedder({"x": q2, "y": (0.6 * qm[1]).astype(np.float32)}))
    raw.audio_database.extend(db2)
    for text, mode, rows, has, qv in (("x", "asr", a2, ha, q2), ("x", "caption", c2, hc, q2),
                                      ("y", "combined", m2, everything, (0.6 * qm[1]).astype(np.float32))):
        idx, sims = no.clean_search(qv, rows, has)
        cos_idx, _ = no.clean_search(qv / np.linalg.norm(qv), rows / np.maximum(np.linalg.norm(rows, axis=1, keepdims=True), 1e-30), has)
        got = raw.search_audio(text, mode)
        assert [int(r["segment_id"][4:]) for r in got] == idx.tolist()
        assert [r["similarity"] for r in got] == sims.tolist()
        if mode == "asr":
            assert idx.tolist() != cos_idx.tolist()                               # the scales really change the ranking

    class ReferenceLike:
        def __init__(self):
            self.text_embedder = FakeEmbedder({"x": q[0]})
            self.audio_database = db

        def search_audio(self, *a, **k):
            raise AssertionError("CPU path must not run")
    patched = legacy.accelerate_clean(ReferenceLike(), dtype=dtype)
    idx, sims = no.clean_search(q[0], a, ha)
    got = patched.search_audio("x", "asr")
    assert [int(r["segment_id"][4:]) for r in got] == idx.tolist() and [r["similarity"] for r in got] == sims.tolist()


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_query_tiles_are_bit_identical_to_one_query_at_a_time(dtype):
    """Several queries per call share one corpus pass (register tiles of 4 / 2 / 1 queries): every
    score must be BIT-identical to the same query scored alone, host and device queries, per-query
    class weights, and a NaN query must poison only its own row of the result."""
    import torch
    rng = np.random.default_rng(5)
    n, nq = 70_001, 11                                   # tiles: 4 + 4 + 3 (padded tile of 4)
    a = rng.standard_normal((n, 384)).astype(np.float32)
    b = rng.standard_normal((n, 384)).astype(np.float32)
    flags = (3 | (rng.integers(0, 4, n).astype(np.uint8) << 2)).astype(np.uint8)
    q = rng.standard_normal((nq, 384)).astype(np.float32)
    cw = rng.uniform(-1, 1, (nq, 4, 2)).astype(np.float32)
    idx = SegmentIndex(dtype, device=0)
    idx.append(a, b, flags)
    alone = np.stack([idx.score_all(q[i], cw[i])[0] for i in range(nq)])
    for m in (2, 3, 4, 5, 11):
        got = idx.score_all(q[:m], cw[:m])
        assert np.array_equal(got, alone[:m]), m
        dev = idx.score_all(torch.from_numpy(q[:m]).cuda(), cw[:m]).cpu().numpy()
        assert np.array_equal(dev, alone[:m]), m
    bad = q[:6].copy()
    bad[4, 100] = np.inf
    dev = idx.score_all(torch.from_numpy(bad).cuda(), cw[:6]).cpu().numpy()
    assert np.isnan(dev[4]).all()
    assert np.array_equal(np.delete(dev, 4, axis=0), np.delete(alone[:6], 4, axis=0))
    with pytest.raises(ValueError):
        idx.score_all(bad, cw[:6])
    assert np.array_equal(idx.score_all(q[:6], cw[:6]), alone[:6])       # nothing sticky
    idx.close()
