"""LIVE pin of the numpy oracle against the unmodified reference (build container only;
skipped where /root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest

from multimodal_audio_search_b200 import synth
from oracle import numpy_oracle as no
from oracle import reference_shim as rs

pytestmark = pytest.mark.skipif(not rs.available(), reason="reference sources not present")


@pytest.mark.parametrize("seed,n,partial", [(101, 150, False), (102, 96, True)])
def test_search_matches_reference(seed, n, partial):
    texts = ["zzz", "piano and singing voice"]
    a, b, f, _ = synth.library(seed, n, len(texts), 8, partial)
    q = synth.raw_queries(seed, 0, len(texts))
    eng = rs.reference_engine(rs.segments_from_arrays(a, b, f), dict(zip(texts, q)))
    for qi, text in enumerate(texts):
        res, wi = eng.search_with_fusion(text)
        wa, wb, an = no.analyze_query_for_weights(text)
        assert (wa, wb, an) == (wi["asr_weight"], wi["audio_weight"], wi["analysis"])
        o = no.search(q[qi], a, b, f, wa, wb)
        assert [int(r["segment_id"][4:]) for r in res] == list(o.indices)
        np.testing.assert_allclose([r["fusion_score"] for r in res], o.fusion, atol=1e-6, rtol=0)
    assert eng.stats["search_pipeline"].total_calls == len(texts)


def test_weights_property():
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    import json, os
    tab = json.load(open(os.path.join(os.path.dirname(rs.__file__), "keyword_table.json")))
    words = list(tab["asr_keywords"]) + list(tab["audio_keywords"]) + ["the", "a", "x", "İ", "ß"]

    @settings(max_examples=150, deadline=None)
    @given(st.lists(st.sampled_from(words), max_size=7), st.booleans())
    def run(ws, upper):
        s = " ".join(ws)
        s = s.upper() if upper else s
        assert no.analyze_query_for_weights(s) == rs.reference_weights(s)
    run()


def test_accelerate_accepts_the_real_reference_object():
    """`accelerate()` on an actual reference DualPipelineAudioSearch instance: the patched method
    uses the reference's own attributes (audio_segments, text_embedder, stats,
    _analyze_query_for_weights).  Without a GPU it must return ([], {}) for an empty library and
    raise -- never compute -- for a populated one."""
    from multimodal_audio_search_b200 import _native as N
    from multimodal_audio_search_b200 import accelerate
    a, b, f, _ = synth.library(5, 12, 1, 4)
    q = synth.raw_queries(5, 0, 1)
    eng = rs.reference_engine([], {"zzz": q[0]})
    original = type(eng).search_with_fusion
    accelerate(eng)
    assert eng.search_with_fusion.__func__ is not original
    assert eng.search_with_fusion("zzz") == ([], {})
    assert eng.stats["search_pipeline"].total_calls == 0
    eng.audio_segments.extend(rs.segments_from_arrays(a, b, f))
    if N.lib().cab_device_count() == 0:
        with pytest.raises(N.CabError):
            eng.search_with_fusion("zzz")
    else:                                            # on a GPU box: same answer as the reference
        ref = rs.reference_engine(rs.segments_from_arrays(a, b, f), {"zzz": q[0]})
        want, _ = ref.search_with_fusion("zzz")
        got, info = eng.search_with_fusion("zzz")
        assert [r["segment_id"] for r in got] == [r["segment_id"] for r in want]
        assert info["analysis"] == "Balanced (no specific keywords detected)"
