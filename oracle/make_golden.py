#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- mints tests/golden/* by executing the UNMODIFIED reference
(/root/reference/audio_search.py via oracle/reference_shim.py) in the build container.

    python oracle/make_golden.py          # rewrites tests/golden/

Fixtures (all outputs come from the reference's own `search_with_fusion` /
`_analyze_query_for_weights`; inputs are regenerated from seeds by
`multimodal_audio_search_b200.synth`, which is integer-exact and machine independent):

  search_cases.json   synthetic libraries (seed, n_rows, plants, partial flags) x query texts
                      -> top-10 segment indices, fusion scores (float64), both similarities,
                      effective + query weights, analysis string.  One case also "peels" the
                      reference 10 results at a time to pin the order of the first 100.
  known_answer.npz    the hand-built 9-segment library of SURVEY.md Appendix B (explicit
                      vectors incl. None / zero / scaled rows) + the reference's answers.
  flag_cases.json     libraries where `*_success` and "embedding present" disagree.
  query_weights.json  query strings -> (asr_weight, audio_weight, analysis).
  clean_search.json   `previous_iterations/clean_audio_search.py:293-320` (`search_audio`): top-10
                      segment ids and similarities for the modes asr / caption / combined.
  legacy_scores.npz   the earlier engine (`previous_iterations/streamlit_app.py:173-223`,
                      `UnifiedAudioSearch.search`): all-N similarity vectors for the strategies
                      asr_only / caption_only / adaptive on seeded libraries.
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multimodal_audio_search_b200 import synth  # noqa: E402
from oracle import numpy_oracle as no  # noqa: E402
from oracle import reference_shim as rs  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

QUERY_TEXTS = [
    "zzz",                                  # balanced 0.5/0.5
    "guitar solo",                          # audio-focused
    "someone speaking clearly",             # mixed
    "what did he say about the drums",      # mixed 0.4/0.6
    "lyrics of the song",                   # ASR-focused
    "loud fast drum beat with heavy bass guitar",
]

SEARCH_CASES = [
    # name, seed, n_rows, plants per query, partial flags
    ("config1_1000", 20261018, 1000, 12, False),
    ("partial_600", 11, 600, 15, True),
    ("tiny_7", 5, 7, 1, True),
    ("noplants_400", 3, 400, 0, False),
    ("odd_333", 99, 333, 9, True),
]


def _result_rows(res):
    return {
        "indices": [int(r["segment_id"][4:]) for r in res],
        "fusion": [float(r["fusion_score"]) for r in res],
        "asr_sim": [float(r["asr_similarity"]) for r in res],
        "audio_sim": [float(r["audio_similarity"]) for r in res],
        "eff_asr_w": [float(r["effective_asr_weight"]) for r in res],
        "eff_audio_w": [float(r["effective_audio_weight"]) for r in res],
    }


def search_cases():
    out = []
    for name, seed, n, plants, partial in SEARCH_CASES:
        nq = len(QUERY_TEXTS)
        a, b, f, _ = synth.library(seed, n, n_queries=nq, plants=plants, partial=partial)
        q = synth.raw_queries(seed, 0, nq)
        eng = rs.reference_engine(rs.segments_from_arrays(a, b, f), dict(zip(QUERY_TEXTS, q)))
        case = {"name": name, "seed": seed, "n_rows": n, "plants": plants, "partial": partial,
                "n_queries": nq, "queries": []}
        for qi, text in enumerate(QUERY_TEXTS):
            res, wi = eng.search_with_fusion(text)
            rec = {"text": text, "qi": qi, "asr_weight": wi["asr_weight"],
                   "audio_weight": wi["audio_weight"], "analysis": wi["analysis"]}
            rec.update(_result_rows(res))
            if res:
                assert set(res[0].keys()) >= {"query_asr_weight", "query_audio_weight"}
            case["queries"].append(rec)
            print(name, text, rec["indices"])
        out.append(case)

    # top-100 order by peeling the reference ten results at a time (segments are scored
    # independently, so removing returned segments does not change the others' scores).
    name, seed, n, plants, partial = ("peel100_500", 77, 500, 150, True)
    a, b, f, _ = synth.library(seed, n, n_queries=1, plants=plants, partial=partial)
    q = synth.raw_queries(seed, 0, 1)
    segs = rs.segments_from_arrays(a, b, f)
    eng = rs.reference_engine(list(segs), {"zzz": q[0]})
    peeled = []
    for _ in range(10):
        res, wi = eng.search_with_fusion("zzz")
        if not res:
            break
        peeled.extend(res)
        drop = {r["segment_id"] for r in res}
        eng.audio_segments = [s for s in eng.audio_segments if s["segment_id"] not in drop]
    rec = {"text": "zzz", "qi": 0, "asr_weight": 0.5, "audio_weight": 0.5,
           "analysis": wi["analysis"] if wi else ""}
    rec.update(_result_rows(peeled))
    out.append({"name": name, "seed": seed, "n_rows": n, "plants": plants, "partial": partial,
                "n_queries": 1, "k": 100, "queries": [rec]})
    print(name, len(peeled))
    with open(os.path.join(GOLD, "search_cases.json"), "w") as fjs:
        json.dump(out, fjs, indent=1)


def _unit_with_cos(q, c, rng):
    """A unit vector with cosine c to unit vector q."""
    g = rng.standard_normal(q.shape[0])
    g -= g.dot(q) * q
    g /= np.linalg.norm(g)
    return (c * q + np.sqrt(1.0 - c * c) * g).astype(np.float32)


def known_answer():
    """SURVEY.md Appendix B library; q = e0."""
    rng = np.random.default_rng(0)
    d = synth.DIM
    q = np.zeros(d, dtype=np.float32)
    q[0] = 1.0
    v = lambda c: _unit_with_cos(q.astype(np.float64), c, rng)  # noqa: E731
    lib = [
        (v(.5), v(.5)), (None, None), (v(.9), None), (None, v(.3)), (v(-.5), v(.9)),
        (v(.1), v(.1)), (3 * v(.7), 0.01 * v(.7)), (np.zeros(d, np.float32), v(.2)),
        (v(-.2), v(-.3)),
    ]
    lib[1] = (lib[0][0].copy(), lib[0][1].copy())          # seg1 == seg0 -> exact tie
    asr = np.zeros((len(lib), d), np.float32)
    aud = np.zeros((len(lib), d), np.float32)
    has_a = np.zeros(len(lib), bool)
    has_b = np.zeros(len(lib), bool)
    segs = []
    for i, (ea, eb) in enumerate(lib):
        if ea is not None:
            asr[i], has_a[i] = ea, True
        if eb is not None:
            aud[i], has_b[i] = eb, True
        # seg7: zero ASR vector but asr_success stays True (the flag follows the text)
        segs.append(rs.make_segment(i, ea, eb))
    flags = (has_a.astype(np.uint8) | (has_b.astype(np.uint8) << 1))
    queries = {"zzz": q, "guitar solo": (2.5 * q).astype(np.float32)}
    eng = rs.reference_engine(segs, queries)
    answers = {}
    for text in queries:
        res, wi = eng.search_with_fusion(text)
        rec = {"asr_weight": wi["asr_weight"], "audio_weight": wi["audio_weight"],
               "analysis": wi["analysis"]}
        rec.update(_result_rows(res))
        answers[text] = rec
        print("known", text, rec["indices"], rec["fusion"])
    # empty library -> ([], {}) before any stats update (audio_search.py:626-627)
    e = rs.reference_engine([], queries)
    assert e.search_with_fusion("zzz") == ([], {})
    assert e.stats["search_pipeline"].total_calls == 0
    np.savez(os.path.join(GOLD, "known_answer.npz"), asr=asr, audio=aud, has_asr=has_a,
             has_audio=has_b, flags=flags, q_zzz=queries["zzz"], q_guitar=queries["guitar solo"],
             answers=json.dumps(answers))


def flag_cases():
    """success flag != embedding present (not produced by the reference's ingest, but legal
    input to `search_with_fusion`): the similarity uses the embedding, the weight uses the flag."""
    seed, n = 1234, 64
    a, b, _, _ = synth.library(seed, n, n_queries=2, plants=10, partial=False)
    q = synth.raw_queries(seed, 0, 2)
    rnd = random.Random(5)
    flags = np.array([rnd.choice([0, 1, 2, 3, 3, 3]) for _ in range(n)], dtype=np.uint8)
    has_a = np.array([rnd.random() < 0.8 for _ in range(n)])
    has_b = np.array([rnd.random() < 0.8 for _ in range(n)])
    segs = rs.segments_from_arrays(a, b, flags, has_a, has_b)
    texts = ["zzz", "guitar solo"]
    eng = rs.reference_engine(segs, dict(zip(texts, q)))
    out = {"seed": seed, "n_rows": n, "plants": 10, "n_queries": 2, "flags": flags.tolist(),
           "has_asr": has_a.astype(int).tolist(), "has_audio": has_b.astype(int).tolist(),
           "queries": []}
    for qi, text in enumerate(texts):
        res, wi = eng.search_with_fusion(text)
        rec = {"text": text, "qi": qi, "asr_weight": wi["asr_weight"],
               "audio_weight": wi["audio_weight"], "analysis": wi["analysis"]}
        rec.update(_result_rows(res))
        out["queries"].append(rec)
        print("flags", text, rec["indices"])
    with open(os.path.join(GOLD, "flag_cases.json"), "w") as fjs:
        json.dump(out, fjs, indent=1)


def query_weights():
    tab = json.load(open(os.path.join(ROOT, "oracle", "keyword_table.json")))
    asr, aud = list(tab["asr_keywords"]), list(tab["audio_keywords"])
    rnd = random.Random(20261018)
    qs = list(QUERY_TEXTS) + [
        "", " ", "conversation about technology", "RHYTHMIC Microphone", "He CALLED me, calling",
        "the beats are powerful piano romantic electronic ambient", "bass", "whispering chorus",
        "hi-hat and cross-rhythm", "acoustic guitar string section", "İstanbul şarkı sözleri",
        "straße MUSIC", "call", "called", "recorded recording", "clear harmony",
    ]
    filler = ["the", "a", "of", "with", "find", "me", "some", "where", "xyz", "42"]
    for _ in range(220):
        na, nb = rnd.choice([(0, 0), (1, 0), (2, 0), (3, 0), (5, 0), (0, 1), (0, 2), (0, 4),
                             (1, 1), (1, 2), (2, 1), (3, 1), (1, 4), (2, 5), (4, 4)])
        words = rnd.sample(asr, na) + rnd.sample(aud, nb) + rnd.sample(filler, rnd.randint(0, 4))
        rnd.shuffle(words)
        s = " ".join(words)
        if rnd.random() < 0.3:
            s = s.title()
        qs.append(s)
    out = []
    for s in qs:
        wa, wb, an = rs.reference_weights(s)
        out.append({"query": s, "asr_weight": wa, "audio_weight": wb, "analysis": an})
    with open(os.path.join(GOLD, "query_weights.json"), "w") as fjs:
        json.dump(out, fjs, indent=0, ensure_ascii=True)
    print("weights", len(out))


LEGACY_CASES = [
    # name, seed, n_rows, plants per query, partial (some embeddings None)
    ("legacy_500", 77, 500, 10, True),
    ("legacy_full_257", 78, 257, 6, False),
]
LEGACY_QUERIES = 2


def legacy_scores():
    out = {}
    meta = []
    for name, seed, n, plants, partial in LEGACY_CASES:
        a, b, f, _ = synth.library(seed, n, LEGACY_QUERIES, plants, partial)
        q = synth.raw_queries(seed, 0, LEGACY_QUERIES)
        has_a, has_b = (f & 1).astype(bool), (f & 2).astype(bool)
        good = no.legacy_good_speech(seed, n)
        db = rs.legacy_database(a, b, has_a, has_b, good)
        for qi in range(LEGACY_QUERIES):
            for strategy in ("asr_only", "caption_only", "adaptive"):
                sims = rs.legacy_search(f"query {qi}", q[qi], db, strategy)
                out[f"{name}/{qi}/{strategy}"] = np.asarray(sims)
        meta.append({"name": name, "seed": seed, "n_rows": n, "plants": plants, "partial": partial,
                     "n_queries": LEGACY_QUERIES})
    np.savez_compressed(os.path.join(GOLD, "legacy_scores.npz"), meta=json.dumps(meta), **out)
    print("legacy", len(out))


CLEAN_CASES = [("clean_700", 55, 700, 14), ("clean_90", 56, 90, 4)]


def clean_cases():
    out = []
    for name, seed, n, plants in CLEAN_CASES:
        a, c, m, ha, hc, q, qm = no.clean_library(seed, n, 2, plants)
        db = rs.clean_database(a, c, m, ha, hc)
        rec = {"name": name, "seed": seed, "n_rows": n, "plants": plants, "queries": []}
        for qi in range(2):
            for mode, qv in (("asr", q[qi]), ("caption", q[qi]), ("combined", qm[qi]), ("no_such_mode", q[qi])):
                res = rs.clean_search(f"q{qi}", qv, db, mode)
                rec["queries"].append({"qi": qi, "mode": mode,
                                       "indices": [int(r["segment_id"][4:]) for r in res],
                                       "similarity": [r["similarity"] for r in res]})
        out.append(rec)
    with open(os.path.join(GOLD, "clean_search.json"), "w") as fjs:
        json.dump(out, fjs, indent=1)
    print("clean", sum(len(r["queries"]) for r in out))


if __name__ == "__main__":
    if not rs.available():
        sys.exit("reference not present: goldens can only be minted in the build container")
    os.makedirs(GOLD, exist_ok=True)
    query_weights()
    known_answer()
    flag_cases()
    search_cases()
    if rs.legacy_available():
        legacy_scores()
    if rs.clean_available():
        clean_cases()
