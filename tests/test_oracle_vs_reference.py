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
