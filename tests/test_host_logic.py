"""CPU-only host logic: the product's query-weight rule against the reference-minted goldens,
and the synthetic generator's structural guarantees."""
import numpy as np

from multimodal_audio_search_b200 import query_weights, synth


def test_query_weights_match_reference(query_weight_cases):
    for rec in query_weight_cases:
        got = query_weights.analyze_query_for_weights(rec["query"])
        assert got == (rec["asr_weight"], rec["audio_weight"], rec["analysis"]), rec["query"]


def test_weight_range_and_sum():
    for a in range(0, 9):
        for b in range(0, 9):
            wa, wb, _ = query_weights.weights_from_matches(a, b)
            assert 0.2 - 1e-12 <= wa <= 0.8 + 1e-12 and abs(wa + wb - 1.0) < 1e-12
    assert query_weights.weights_from_matches(0, 2)[0] == 0.30000000000000004


def test_synth_shards_are_the_same_library():
    seed, n = 9, 1000
    a, b, f, spec = synth.library(seed, n, n_queries=3, plants=7, partial=True)
    parts = [synth.library(seed, n, 3, 7, True, r0=r0, r1=r1) for r0, r1 in ((0, 333), (333, 900), (900, 1000))]
    np.testing.assert_array_equal(np.concatenate([p[0] for p in parts]), a)
    np.testing.assert_array_equal(np.concatenate([p[1] for p in parts]), b)
    np.testing.assert_array_equal(np.concatenate([p[2] for p in parts]), f)
    assert a.dtype == np.float32 and np.all(a == np.round(a))            # integer valued
    rows = np.concatenate([synth.plant_rows_of_query(spec, q) for q in range(3)])
    assert len(set(rows.tolist())) == 21                                   # collision free
    g = synth.plant_of_rows(spec, np.arange(n))
    assert sorted(g[g >= 0].tolist()) == list(range(21))


def test_synth_plants_are_near_their_query():
    seed, n = 4, 5000
    a, b, _, spec = synth.library(seed, n, n_queries=2, plants=30)
    q = synth.raw_queries(seed, 0, 2)
    for qi in range(2):
        rows = synth.plant_rows_of_query(spec, qi)
        qa = q[qi] / np.linalg.norm(q[qi])
        planted = []
        for t, r in enumerate(rows):
            kind = (qi * 30 + t) % 3
            ca = a[r] @ qa / np.linalg.norm(a[r])
            cb = b[r] @ qa / np.linalg.norm(b[r])
            planted += [ca] if kind == 0 else [cb] if kind == 1 else [ca, cb]
        # cosine ~= m/sqrt(m^2+n^2) with m, n in 1..8: spread over (0.12, 0.99)
        assert min(planted) > -0.05 and max(planted) > 0.9 and np.mean(planted) > 0.5


def test_synth_row_lists_equal_row_ranges():
    """`library(rows=[...])` (what bench.py's parity check regenerates) == slices of the range form."""
    for mode in ("planted", "ascending", "clustered"):
        a, b, f, _ = synth.library(11, 4000, n_queries=3, plants=9, partial=True, mode=mode)
        pick = [3999, 0, 17, 2048, 1234]
        a2, b2, f2, _ = synth.library(11, 4000, n_queries=3, plants=9, partial=True, mode=mode, rows=pick)
        np.testing.assert_array_equal(a[pick], a2)
        np.testing.assert_array_equal(b[pick], b2)
        np.testing.assert_array_equal(f[pick], f2)
        lo, _, _, _ = synth.library(11, 4000, 3, 9, True, r0=1000, r1=1500, mode=mode)
        np.testing.assert_array_equal(lo, a[1000:1500])


def test_synth_stress_distributions_have_the_advertised_shape():
    """ascending: for a query near the ascent direction every row scores above its predecessor
    (the adversarial order for top-k pruning); clustered: a query near a centre has ~N/1024 rows
    far above the 0.1 threshold and the rest of the library far below them."""
    n = 200_000
    unit = lambda x: x / np.linalg.norm(x, axis=-1, keepdims=True)
    a, b, _, _ = synth.library(3, n, mode="ascending", r0=n - 5000, r1=n)
    q = unit(synth.bench_queries(3, "ascending", 0, 4))
    for corpus in (a, b):
        s = unit(corpus) @ q.T
        assert (np.diff(s, axis=0) > 0).mean() > 0.99 and s.min() > 0.5
    rows = np.arange(n)
    cl = synth.cluster_of_rows(3, rows)
    assert cl.min() == 0 and cl.max() == synth.N_CLUSTERS - 1
    q = unit(synth.bench_queries(3, "clustered", 0, 2))
    for qi in range(2):
        mine = rows[cl == qi]
        assert 100 < len(mine) < 320
        a, b, _, _ = synth.library(3, n, mode="clustered", rows=mine)
        assert (unit(a) @ q[qi]).min() > 0.45 and (unit(b) @ q[qi]).min() > 0.45
        a, b, _, _ = synth.library(3, n, mode="clustered", rows=rows[cl != qi][:20000])
        assert np.abs(unit(a) @ q[qi]).max() < 0.3


def test_bucketed_keyword_count_equals_the_literal_form():
    """query_weights.count_matches goes through first-bigram buckets; it must equal the literal
    `sum(n for kw in table if kw in query)` on the golden queries, on every keyword itself, on
    concatenations and on random strings over the keywords' alphabet."""
    from multimodal_audio_search_b200.keyword_table import ASR_KEYWORDS, AUDIO_KEYWORDS
    rng = np.random.default_rng(0)
    kws = list(ASR_KEYWORDS) + list(AUDIO_KEYWORDS)
    alphabet = sorted(set("".join(kws)) | {" "})
    cases = []
    cases += kws + ["", "a", "xy"] + [kws[i] + kws[-i - 1] for i in range(0, len(kws), 7)]
    cases += [" ".join(rng.choice(kws, 4)) for _ in range(200)]
    cases += ["".join(rng.choice(alphabet, rng.integers(1, 40))) for _ in range(2000)]
    for q in cases:
        assert query_weights.count_matches(q) == query_weights.count_matches_plain(q), q
