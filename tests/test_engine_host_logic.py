"""CPU: the host logic of the drop-in engine (`engine.py`) -- library sync, result dictionaries,
weight_info, stats, the columnar table path, `search_many` -- with the device index replaced by an
oracle-backed fake.  Expected values are the golden outputs minted from the reference
(tests/golden/search_cases.json), so what is checked is exactly the Python around the GPU call;
the GPU run of the same checks is tests/test_gpu_engine.py."""
import numpy as np
import pytest

from multimodal_audio_search_b200 import engine, synth
from multimodal_audio_search_b200.index import SearchResult
from multimodal_audio_search_b200.segment_table import SegmentRecord, SegmentTable
from oracle import numpy_oracle as no
from oracle.reference_shim import FakeEmbedder, segments_from_arrays


class FakeSegmentIndex:
    """SegmentIndex look-alike backed by the numpy oracle (test infrastructure only)."""
    instances = []

    def __init__(self, dtype="fp32", capacity=0, device=0):
        self.dtype, self.device, self.row_base = dtype, device, 0
        self.a = np.zeros((0, 384), np.float32)
        self.b = np.zeros((0, 384), np.float32)
        self.f = np.zeros(0, np.uint8)
        self.appends, self.clears = 0, 0
        FakeSegmentIndex.instances.append(self)

    def __len__(self):
        return len(self.f)

    def append(self, asr, audio, flags):
        for m in (asr, audio):
            if not np.isfinite(m).all():
                raise ValueError("Input contains NaN or infinity (rows not appended)")
        self.a, self.b = np.vstack([self.a, asr]), np.vstack([self.b, audio])
        self.f = np.concatenate([self.f, flags])
        self.appends += 1

    def clear(self):
        self.a, self.b, self.f = self.a[:0], self.b[:0], self.f[:0]
        self.clears += 1

    def read_rows(self, corpus, r0, r1):
        return no.normalize_rows((self.a, self.b)[corpus][r0:r1])

    def search(self, queries, w_asr=0.5, w_audio=0.5, k=10, threshold=0.1, path="auto"):
        q = np.atleast_2d(queries)
        wa = np.broadcast_to(np.asarray(w_asr, np.float64), (len(q),))
        wb = np.broadcast_to(np.asarray(w_audio, np.float64), (len(q),))
        out = SearchResult(np.full((len(q), k), -1, np.int64), np.zeros((len(q), k)), np.zeros((len(q), k), np.float32),
                           np.zeros((len(q), k), np.float32), np.zeros((len(q), k), np.uint8), np.zeros(len(q), np.int32))
        for i in range(len(q)):
            o = no.search(q[i], self.a, self.b, self.f, wa[i], wb[i], k=k, threshold=threshold,
                          has_asr=np.any(self.a != 0, axis=1), has_audio=np.any(self.b != 0, axis=1))
            c = len(o.indices)
            out.indices[i, :c], out.fusion[i, :c], out.count[i] = o.indices, o.fusion, c
            out.asr_sim[i, :c], out.audio_sim[i, :c], out.flags[i, :c] = o.asr_sim, o.audio_sim, self.f[o.indices]
        return out


@pytest.fixture(autouse=True)
def fake_index(monkeypatch):
    FakeSegmentIndex.instances = []
    monkeypatch.setattr(engine, "SegmentIndex", FakeSegmentIndex)


def _case_engine(case):
    a, b, f, _ = synth.library(case["seed"], case["n_rows"], case["n_queries"], case["plants"], case["partial"])
    q = synth.raw_queries(case["seed"], 0, case["n_queries"])
    eng = engine.DualPipelineAudioSearch(text_embedder=FakeEmbedder({r["text"]: q[r["qi"]] for r in case["queries"]}))
    return eng, segments_from_arrays(a, b, f)


def _check(results, info, rec):
    assert [int(r["segment_id"][4:]) for r in results] == rec["indices"]
    assert (info["asr_weight"], info["audio_weight"], info["analysis"], info["query"]) == \
        (rec["asr_weight"], rec["audio_weight"], rec["analysis"], rec["text"])
    for j, r in enumerate(results):
        assert abs(r["fusion_score"] - rec["fusion"][j]) <= 1e-6
        assert r["effective_asr_weight"] == rec["eff_asr_w"][j] and r["effective_audio_weight"] == rec["eff_audio_w"][j]
        assert r["query_asr_weight"] == rec["asr_weight"] and r["query_audio_weight"] == rec["audio_weight"]
        assert all(type(r[k]) is float for k in ("asr_similarity", "audio_similarity", "fusion_score"))


def test_results_weight_info_and_stats_match_the_reference(search_cases):
    for case in search_cases[:3]:
        if case.get("k", 10) != 10:
            continue
        eng, segs = _case_engine(case)
        assert eng.search_with_fusion("zzz") == ([], {}) and eng.stats["search_pipeline"].total_calls == 0   # :626-627
        eng.audio_segments.extend(segs)
        hits = 0
        for rec in case["queries"]:
            results, info = eng.search_with_fusion(rec["text"])
            _check(results, info, rec)
            assert all(list(r.keys())[:12] == list(segs[0].keys()) for r in results)     # {**segment, ...} key order
            hits += bool(results)
        st = eng.stats["search_pipeline"]
        assert st.total_calls == len(case["queries"]) and st.successful_extractions == hits
        assert st.success_rate == hits / len(case["queries"]) and st.avg_processing_time == st.total_processing_time / st.total_calls


def test_library_sync_is_incremental_and_rebuilds_when_replaced(search_cases):
    case = search_cases[1]
    eng, segs = _case_engine(case)
    text = case["queries"][0]["text"]
    eng.audio_segments.extend(segs[:100])
    eng.search_with_fusion(text)
    idx = FakeSegmentIndex.instances[-1]
    assert (len(idx), idx.appends, idx.clears) == (100, 1, 0)
    eng.search_with_fusion(text)
    assert idx.appends == 1                                          # nothing new: no device work
    eng.audio_segments.extend(segs[100:])                            # :797
    results, info = eng.search_with_fusion(text)
    assert (len(idx), idx.appends, idx.clears) == (len(segs), 2, 0)
    _check(results, info, case["queries"][0])
    eng.audio_segments = list(segs[:50])                             # replaced / shrunk list: rebuild
    eng.search_with_fusion(text)
    assert (len(idx), idx.clears) == (50, 1)
    bad = dict(segs[0]); bad["asr_embedding"] = np.full(384, np.nan, np.float32)
    eng.audio_segments.append(bad)
    with pytest.raises(ValueError, match="NaN or infinity"):
        eng.search_with_fusion(text)
    wrong = dict(segs[0]); wrong["audio_embedding"] = np.zeros(100, np.float32)
    eng.audio_segments[-1] = wrong
    with pytest.raises(ValueError, match="Incompatible dimension"):
        eng.search_with_fusion(text)


def test_columnar_table_path_and_search_many(search_cases):
    case = search_cases[1]
    eng, segs = _case_engine(case)
    eng.audio_segments = SegmentTable()
    eng.audio_segments.extend(segs, file="a.wav")
    texts = [r["text"] for r in case["queries"]]
    many = eng.search_many(texts)
    assert eng.audio_segments.n_pending == 0                         # embeddings handed over once
    for (results, info), rec in zip(many, case["queries"]):
        _check(results, info, rec)
        assert all(isinstance(r, SegmentRecord) and r["file"] == "a.wav" for r in results)
        for r in results:                                            # lazy embedding = the index's normalised row
            i = int(r["segment_id"][4:])
            e = r["asr_embedding"]
            assert (e is None) == (segs[i]["asr_embedding"] is None)
            if e is not None:
                assert np.allclose(e, segs[i]["asr_embedding"] / np.linalg.norm(segs[i]["asr_embedding"]), atol=1e-6)
    single = [eng.search_with_fusion(t) for t in texts]
    assert [[r["segment_id"] for r in rs] for rs, _ in single] == [[r["segment_id"] for r in rs] for rs, _ in many]
    assert eng.stats["search_pipeline"].total_calls == 2 * len(texts)
    assert eng.search_many([]) == []
    # a table that is out of step with the index is refused, not searched
    other = SegmentTable.from_columns(5)
    eng.audio_segments = other
    with pytest.raises(RuntimeError, match="rows"):
        eng.search_with_fusion(texts[0])


def test_accelerate_keeps_the_objects_own_helpers(search_cases):
    case = search_cases[0]
    a, b, f, _ = synth.library(case["seed"], case["n_rows"], case["n_queries"], case["plants"], case["partial"])
    q = synth.raw_queries(case["seed"], 0, case["n_queries"])
    calls = []

    class ReferenceLike:
        def __init__(self):
            self.audio_segments = segments_from_arrays(a, b, f)
            self.text_embedder = FakeEmbedder({r["text"]: q[r["qi"]] for r in case["queries"]})
            self.stats = {"search_pipeline": engine.PipelineStats("Search Pipeline", "Cosine Similarity")}

        def _analyze_query_for_weights(self, query):
            calls.append(query)
            return engine.query_weights.analyze_query_for_weights(query)

        def search_with_fusion(self, query):
            raise AssertionError("CPU path must not run")
    eng = engine.accelerate(ReferenceLike(), columnar=True)
    assert isinstance(eng.audio_segments, SegmentTable)
    rec = case["queries"][0]
    results, info = eng.search_with_fusion(rec["text"])
    _check(results, info, rec)
    assert calls == [rec["text"]]
