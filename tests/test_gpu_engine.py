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
