"""GPU: the drop-in engine (`DualPipelineAudioSearch`, `accelerate`) reproduces the reference's
`search_with_fusion` return value (golden fixtures minted from the reference)."""
import numpy as np
import pytest

from multimodal_audio_search_b200 import DualPipelineAudioSearch, accelerate, synth
from oracle.reference_shim import FakeEmbedder, segments_from_arrays

pytestmark = pytest.mark.gpu

RESULT_KEYS = {"segment_id", "start_time", "end_time", "duration", "asr_text", "asr_embedding",
               "asr_success", "audio_description", "audio_embedding", "audio_success", "audio_data",
               "sample_rate", "asr_similarity", "audio_similarity", "fusion_score",
               "effective_asr_weight", "effective_audio_weight", "query_asr_weight", "query_audio_weight"}


def _check(results, info, rec):
    assert [int(r["segment_id"][4:]) for r in results] == rec["indices"]
    assert (info["asr_weight"], info["audio_weight"], info["analysis"], info["query"]) == \
        (rec["asr_weight"], rec["audio_weight"], rec["analysis"], rec["text"])
    for j, r in enumerate(results):
        assert set(r.keys()) == RESULT_KEYS
        assert all(type(r[k]) is float for k in ("asr_similarity", "audio_similarity", "fusion_score",
                                                 "effective_asr_weight", "effective_audio_weight"))
        assert abs(r["fusion_score"] - rec["fusion"][j]) <= 1e-5
        assert abs(r["asr_similarity"] - rec["asr_sim"][j]) <= 1e-5
        assert abs(r["audio_similarity"] - rec["audio_sim"][j]) <= 1e-5
        assert r["effective_asr_weight"] == rec["eff_asr_w"][j]
        assert r["effective_audio_weight"] == rec["eff_audio_w"][j]
        assert r["query_asr_weight"] == rec["asr_weight"] and r["query_audio_weight"] == rec["audio_weight"]


def test_standalone_engine_matches_reference_outputs(search_cases):
    for case in search_cases:
        if case.get("k", 10) != 10:
            continue
        a, b, f, _ = synth.library(case["seed"], case["n_rows"], case["n_queries"], case["plants"], case["partial"])
        q = synth.raw_queries(case["seed"], 0, case["n_queries"])
        eng = DualPipelineAudioSearch(text_embedder=FakeEmbedder({r["text"]: q[r["qi"]] for r in case["queries"]}))
        assert eng.search_with_fusion("zzz") == ([], {}) and eng.stats["search_pipeline"].total_calls == 0
        segs = segments_from_arrays(a, b, f)
        half = len(segs) // 2
        eng.audio_segments.extend(segs[:half])            # the reference's only mutation (:797)
        eng.search_with_fusion("zzz")
        eng.audio_segments.extend(segs[half:])            # incremental sync
        calls = eng.stats["search_pipeline"].total_calls
        for rec in case["queries"]:
            results, info = eng.search_with_fusion(rec["text"])
            _check(results, info, rec)
        st = eng.stats["search_pipeline"]
        assert st.total_calls == calls + len(case["queries"])
        n_hit = sum(1 for r in case["queries"] if r["indices"])
        assert st.successful_extractions >= n_hit


def test_accelerate_patches_a_reference_like_object(search_cases):
    case = search_cases[1]
    a, b, f, _ = synth.library(case["seed"], case["n_rows"], case["n_queries"], case["plants"], case["partial"])
    q = synth.raw_queries(case["seed"], 0, case["n_queries"])

    class ReferenceLike:                                    # the attributes the patch relies on
        def __init__(self):
            from multimodal_audio_search_b200 import PipelineStats, analyze_query_for_weights
            self.audio_segments = segments_from_arrays(a, b, f)
            self.text_embedder = FakeEmbedder({r["text"]: q[r["qi"]] for r in case["queries"]})
            self.stats = {"search_pipeline": PipelineStats("Search Pipeline", "Cosine Similarity")}
            self._analyze_query_for_weights = analyze_query_for_weights

        def search_with_fusion(self, query):
            raise AssertionError("CPU path must not run")
    eng = accelerate(ReferenceLike())
    for rec in case["queries"]:
        results, info = eng.search_with_fusion(rec["text"])
        _check(results, info, rec)


def test_replaced_library_is_resynced():
    a, b, f, _ = synth.library(3, 50, 1, 10)
    q = synth.raw_queries(3, 0, 1)
    eng = DualPipelineAudioSearch(text_embedder=FakeEmbedder({"zzz": q[0]}))
    eng.audio_segments = segments_from_arrays(a, b, f)
    r1, _ = eng.search_with_fusion("zzz")
    eng.audio_segments = segments_from_arrays(a[::-1].copy(), b[::-1].copy(), f[::-1].copy())
    r2, _ = eng.search_with_fusion("zzz")
    assert [49 - int(r["segment_id"][4:]) for r in r2] == [int(r["segment_id"][4:]) for r in r1]
    with pytest.raises(ValueError):
        eng.audio_segments.append({**eng.audio_segments[0], "asr_embedding": np.ones(100, np.float32)})
        eng.search_with_fusion("zzz")


def test_columnar_library_gives_the_same_results_and_survives_a_session(search_cases, tmp_path):
    """SegmentTable instead of the list of dicts (8(f) rank 3) + save/load (rank 1): same
    `search_with_fusion` return value; lazy audio_data / embeddings resolve to the right rows."""
    from multimodal_audio_search_b200 import SegmentRecord, SegmentTable
    case = search_cases[1]
    a, b, f, _ = synth.library(case["seed"], case["n_rows"], case["n_queries"], case["plants"], case["partial"])
    q = synth.raw_queries(case["seed"], 0, case["n_queries"])
    embedder = FakeEmbedder({r["text"]: q[r["qi"]] for r in case["queries"]})
    segs = segments_from_arrays(a, b, f)
    for i, s in enumerate(segs):
        s["audio_data"] = np.full(8, i, dtype=np.float32)

    eng = DualPipelineAudioSearch(text_embedder=embedder)
    eng.audio_segments = SegmentTable()
    half = len(segs) // 2
    eng.audio_segments.extend(segs[:half], file="a.wav")
    eng.search_with_fusion(case["queries"][0]["text"])
    eng.audio_segments.extend(segs[half:], file="b.wav")
    assert eng.audio_segments.n_pending == len(segs) - half
    for rec in case["queries"]:
        results, info = eng.search_with_fusion(rec["text"])
        assert all(isinstance(r, SegmentRecord) for r in results)
        _check([{k: v for k, v in r.items() if k != "file"} for r in results], info, rec)
        for r in results:
            row = int(r["segment_id"][4:])
            assert r["file"] == ("a.wav" if row < half else "b.wav")
            assert r["audio_data"][0] == row
            for key, src in (("asr_embedding", a), ("audio_embedding", b)):
                e = r[key]                                    # read back from HBM, L2-normalised
                if not np.any(src[row]):
                    assert e is None
                else:
                    assert np.allclose(e, src[row] / np.linalg.norm(src[row]), atol=1e-6)
    assert eng.audio_segments.n_pending == 0

    path = str(tmp_path / "library.cab")
    eng.save_library(path)
    fresh = DualPipelineAudioSearch(text_embedder=embedder)
    fresh.load_library(path)
    assert len(fresh.audio_segments) == len(segs)
    for rec in case["queries"]:
        results, info = fresh.search_with_fusion(rec["text"])
        _check([{k: v for k, v in r.items() if k != "file"} for r in results], info, rec)
    extra = segments_from_arrays(a[:3], b[:3], f[:3])
    fresh.audio_segments.extend(extra)                        # a loaded library keeps growing
    fresh.search_with_fusion(case["queries"][0]["text"])
    assert len(fresh._cab_library.index) == len(segs) + 3

    # a list-based engine can be saved too, and accelerate(columnar=True) converts in place
    listy = DualPipelineAudioSearch(text_embedder=embedder)
    listy.audio_segments.extend(segs)
    listy.save_library(str(tmp_path / "listy.cab"))
    again = DualPipelineAudioSearch(text_embedder=embedder)
    again.load_library(str(tmp_path / "listy.cab"))
    rec = case["queries"][0]
    results, info = again.search_with_fusion(rec["text"])
    _check([{k: v for k, v in r.items() if k != "file"} for r in results], info, rec)
    conv = accelerate(listy, columnar=True)
    assert isinstance(conv.audio_segments, SegmentTable)
    results, info = conv.search_with_fusion(rec["text"])
    _check([{k: v for k, v in r.items() if k != "file"} for r in results], info, rec)


def test_search_many_and_session_batching(search_cases):
    """Beyond the reference: `search_many` (one scan batch for several query strings) and
    concurrent `search_with_fusion` calls coalesced by a SearchBatcher must return exactly what
    one `search_with_fusion` call per query returns."""
    import threading
    case = search_cases[1]
    a, b, f, _ = synth.library(case["seed"], case["n_rows"], case["n_queries"], case["plants"], case["partial"])
    q = synth.raw_queries(case["seed"], 0, case["n_queries"])
    embedder = FakeEmbedder({r["text"]: q[r["qi"]] for r in case["queries"]})
    eng = DualPipelineAudioSearch(text_embedder=embedder)
    assert eng.search_many(["zzz"]) == [([], {})]
    eng.audio_segments.extend(segments_from_arrays(a, b, f))
    texts = [r["text"] for r in case["queries"]]
    single = [eng.search_with_fusion(t) for t in texts]
    calls = eng.stats["search_pipeline"].total_calls
    many = eng.search_many(texts)
    assert eng.stats["search_pipeline"].total_calls == calls + len(texts)
    for (r1, i1), (r2, i2), rec in zip(single, many, case["queries"]):
        _check(r2, i2, rec)
        assert i1 == i2 and [x["segment_id"] for x in r1] == [x["segment_id"] for x in r2]
        assert [x["fusion_score"] for x in r1] == [x["fusion_score"] for x in r2]

    batcher = eng.enable_batching(max_batch=64, max_wait_s=0.01)
    out = {}

    def session(t):
        out[t] = eng.search_with_fusion(texts[t % len(texts)])
    threads = [threading.Thread(target=session, args=(t,)) for t in range(32)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert batcher.stats.requests == 32 and batcher.stats.batches < 32
    for t, (results, info) in out.items():
        _check(results, info, case["queries"][t % len(texts)])
    more = segments_from_arrays(a[:5], b[:5], f[:5])              # ingest while batching: appended between two batches
    eng.audio_segments.extend(more)
    results, info = eng.search_with_fusion(texts[0])
    assert len(eng._cab_library.index) == len(f) + 5 and info["query"] == texts[0]
    batcher.close()


def test_fp32_engine_with_tensor_core_batches(search_cases):
    """`tensor_core_batches=True`: the fp32 drop-in keeps bf16 shadows; `search_many` with >= 64 query
    strings is preselected on the tensor cores and still returns exactly the reference-golden dicts."""
    case = search_cases[1]
    a, b, f, _ = synth.library(case["seed"], case["n_rows"], case["n_queries"], case["plants"], case["partial"])
    q = synth.raw_queries(case["seed"], 0, case["n_queries"])
    embedder = FakeEmbedder({r["text"]: q[r["qi"]] for r in case["queries"]})
    eng = DualPipelineAudioSearch(text_embedder=embedder, tensor_core_batches=True)
    eng.audio_segments.extend(segments_from_arrays(a, b, f))
    texts = [r["text"] for r in case["queries"]]
    many = eng.search_many([texts[i % len(texts)] for i in range(80)])
    index = eng._cab_library.index
    assert index.get_option("tensor_core_shadow") == 1 and index.get_option("total_shadow_queries") == 80
    for i, (results, info) in enumerate(many):
        _check(results, info, case["queries"][i % len(texts)])
