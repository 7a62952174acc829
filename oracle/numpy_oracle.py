"""TEST INFRASTRUCTURE -- CPU restatement of the reference's search hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this; the product
package never does (it has no CPU fallback).

Restates, vectorised over segments, what `/root/reference/audio_search.py` computes per segment:

  cosine x2          audio_search.py:639-651  via scikit-learn 1.9.0 `cosine_similarity`
                     (sklearn/metrics/pairwise.py:1744-1750: normalize(X), normalize(Y), X @ Y.T;
                     row_norms = sqrt(einsum('ij,ij->i')) in the input dtype (fp32); a zero norm
                     divides by 1 (preprocessing/_data.py:2077-2081); non-finite input raises
                     ValueError).  `None` embedding -> similarity 0.0 (:640-641).
  fusion             audio_search.py:654-670  float64 arithmetic on the widened fp32 sims;
                     effective weight = query weight if the pipeline's success flag else 0,
                     renormalised by their sum; rows with no successful pipeline are skipped.
  threshold          audio_search.py:672      strict `fusion > 0.1` in float64.
  order + top-k      audio_search.py:685,699  stable descending sort => (score desc, index asc).
  query weights      audio_search.py:457-622  substring counting + 4-branch rule.

PARITY PIN: the reference ships no tests or golden vectors for this path ("parity unpinned" by
the reference itself).  This restatement is pinned instead against outputs
of the reference's own code executed in the build container (`oracle/reference_shim.py`):
`tests/test_oracle_vs_reference.py` (live, container only) and the committed fixtures under
`tests/golden/` minted by `oracle/make_golden.py`.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_TABLE = None


def _tables():
    global _TABLE
    if _TABLE is None:
        with open(os.path.join(_HERE, "keyword_table.json")) as f:
            _TABLE = json.load(f)
    return _TABLE


def analyze_query_for_weights(query: str):
    """audio_search.py:457-622.  Returns (asr_weight, audio_weight, analysis)."""
    q = query.lower()                                                     # :459
    t = _tables()
    asr_matches = sum(n for kw, n in t["asr_keywords"].items() if kw in q)      # :586
    audio_matches = sum(n for kw, n in t["audio_keywords"].items() if kw in q)  # :587
    if asr_matches == 0 and audio_matches == 0:                           # :593-596
        return 0.5, 0.5, "Balanced (no specific keywords detected)"
    if asr_matches > 0 and audio_matches == 0:                            # :598-603
        strength = min(asr_matches / 3.0, 1.0)
        asr_weight = 0.5 + (0.3 * strength)
        return asr_weight, 1.0 - asr_weight, f"ASR-focused ({asr_matches} speech keywords)"
    if audio_matches > 0 and asr_matches == 0:                            # :605-610
        strength = min(audio_matches / 3.0, 1.0)
        audio_weight = 0.5 + (0.3 * strength)
        return 1.0 - audio_weight, audio_weight, f"Audio-focused ({audio_matches} audio keywords)"
    total = asr_matches + audio_matches                                   # :612-620
    asr_weight = 0.2 + ((asr_matches / total) * 0.6)
    return asr_weight, 1.0 - asr_weight, f"Mixed query (ASR:{asr_matches}, Audio:{audio_matches})"


def normalize_rows(x: np.ndarray) -> np.ndarray:
    """sklearn `normalize(X, norm='l2')` on fp32 rows; zero rows stay zero; non-finite raises."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if not np.isfinite(x).all():
        raise ValueError("Input contains NaN or infinity.")
    norms = np.sqrt(np.einsum("ij,ij->i", x, x))
    norms[norms == 0.0] = 1.0
    return x / norms[:, None]


def cosine_rows(q: np.ndarray, rows: np.ndarray, has: np.ndarray | None = None) -> np.ndarray:
    """fp32 cosine of one query against every row; rows with has==False give exactly 0.0."""
    qn = normalize_rows(np.asarray(q, dtype=np.float32).reshape(1, -1))[0]
    s = normalize_rows(rows) @ qn
    if has is not None:
        s = np.where(has, s, np.float32(0.0))
    return s.astype(np.float32)


@dataclass
class OracleResult:
    indices: np.ndarray          # int64 (<=k,)  segment indices, best first
    fusion: np.ndarray           # float64
    asr_sim: np.ndarray          # float32
    audio_sim: np.ndarray        # float32
    eff_asr_w: np.ndarray        # float64
    eff_audio_w: np.ndarray      # float64
    all_fusion: np.ndarray       # float64 (N,) fused score of every row (-inf where skipped)


def fuse(s_asr, s_audio, flags, w_asr: float, w_audio: float):
    """audio_search.py:654-670 for every row.  Returns (fusion f64 with -inf where the row is
    skipped, eff_asr_w, eff_audio_w)."""
    sa = np.asarray(s_asr, dtype=np.float32).astype(np.float64)
    sb = np.asarray(s_audio, dtype=np.float32).astype(np.float64)
    flags = np.asarray(flags)
    ea = np.where(flags & 1, np.float64(w_asr), 0.0)
    eb = np.where(flags & 2, np.float64(w_audio), 0.0)
    tot = ea + eb
    ok = (tot > 0) & ((sa > 0) | (sb > 0))                                # :654, :661
    safe = np.where(tot > 0, tot, 1.0)
    ea = ea / safe
    eb = eb / safe
    fusion = ea * sa + eb * sb                                            # :667-670 (mul, mul, add)
    return np.where(ok, fusion, -np.inf), ea, eb


def search(q, asr_rows, audio_rows, flags, w_asr: float, w_audio: float, k: int = 10,
           threshold: float = 0.1, has_asr=None, has_audio=None) -> OracleResult:
    """One query against a library.  `flags` bit0/bit1 = asr_success/audio_success; `has_*` says
    whether the embedding is present (defaults to the success flag, audio_search.py:282-289)."""
    flags = np.asarray(flags, dtype=np.uint8)
    has_asr = (flags & 1).astype(bool) if has_asr is None else np.asarray(has_asr, dtype=bool)
    has_audio = (flags & 2).astype(bool) if has_audio is None else np.asarray(has_audio, dtype=bool)
    sa = cosine_rows(q, asr_rows, has_asr)
    sb = cosine_rows(q, audio_rows, has_audio)
    fusion, ea, eb = fuse(sa, sb, flags, w_asr, w_audio)
    passing = np.nonzero(fusion > threshold)[0]                           # :672 strict, float64
    order = passing[np.argsort(-fusion[passing], kind="stable")][:k]      # :685 stable, :699
    return OracleResult(order.astype(np.int64), fusion[order], sa[order], sb[order], ea[order],
                        eb[order], fusion)


def search_chunked(q, row_source, n_rows: int, w_asr: float, w_audio: float, k: int = 10,
                   threshold: float = 0.1, chunk: int = 65536):
    """Same as `search` for libraries too large to hold: `row_source(r0, r1)` yields
    (asr_rows, audio_rows, flags) for global rows [r0, r1).  Keeps the running best k."""
    best = []
    for r0 in range(0, n_rows, chunk):
        r1 = min(n_rows, r0 + chunk)
        a, b, f = row_source(r0, r1)
        res = search(q, a, b, f, w_asr, w_audio, k, threshold)
        best.extend(zip(-res.fusion, res.indices + r0, res.asr_sim, res.audio_sim))
        best.sort(key=lambda t: (t[0], t[1]))
        best = best[:k]
    idx = np.array([t[1] for t in best], dtype=np.int64)
    return (idx, np.array([-t[0] for t in best], dtype=np.float64),
            np.array([t[2] for t in best], dtype=np.float32),
            np.array([t[3] for t in best], dtype=np.float32))


def search_prenormalized(qn, asr_n, audio_n, flags, w_asr: float, w_audio: float, k: int = 10,
                         threshold: float = 0.1):
    """CPU-baseline variant used by bench.py: rows L2-normalised ONCE at ingest (what the engine
    does) instead of on every call like the reference; per query 2 sgemv + fusion + threshold +
    top-k (argpartition, then a stable sort of the k survivors).  Same results as `search`."""
    sa = asr_n @ qn
    sb = audio_n @ qn
    fusion, _, _ = fuse(sa, sb, flags, w_asr, w_audio)
    passing = np.nonzero(fusion > threshold)[0]
    if len(passing) > k:
        part = passing[np.argpartition(-fusion[passing], k - 1)[:k]]
        kth = fusion[part].min()
        passing = passing[fusion[passing] >= kth]
    order = passing[np.argsort(-fusion[passing], kind="stable")][:k]
    return order.astype(np.int64), fusion[order], sa[order], sb[order]


# ---- the earlier engine's modes (previous_iterations/streamlit_app.py:173-223) --------------------
LEGACY_CLASS_WEIGHTS = {
    # strategy -> {w_asr, w_caption} per row weight class; class 1 = len(transcript.strip()) > 10 (:216)
    "asr_only": ((1.0, 0.0),) * 4,                                        # :188-193
    "caption_only": ((0.0, 1.0),) * 4,                                    # :195-200
    "adaptive": ((0.2, 0.8), (0.7, 0.3), (0.2, 0.8), (0.2, 0.8)),         # :217 / :219
}


def legacy_good_speech(seed: int, n: int) -> np.ndarray:
    """Synthetic libraries: which items carry a transcript longer than 10 characters (:216).
    Integer hash, so fixtures and tests regenerate it from the seed."""
    i = np.arange(n, dtype=np.uint64)
    h = (i * np.uint64(2654435761) + np.uint64(seed) * np.uint64(40503)) & np.uint64(0xFFFFFFFF)
    return ((h >> np.uint64(13)) & np.uint64(1)).astype(bool)


def legacy_scores(q, asr_rows, caption_rows, has_asr, has_caption, row_class, strategy: str = "adaptive") -> np.ndarray:
    """`UnifiedAudioSearch.search` for every item: float32 (N,) similarities, no threshold.
    A `None` embedding gives 0.0 (:203-210).  `w * sim` multiplies a Python float with a
    numpy float32 scalar: float32 arithmetic under NumPy >= 2 promotion (the weight is rounded to
    fp32, each product and the sum are rounded separately); the single-corpus modes return the
    cosine itself."""
    sa = cosine_rows(q, asr_rows, np.asarray(has_asr, dtype=bool))
    sb = cosine_rows(q, caption_rows, np.asarray(has_caption, dtype=bool))
    if strategy == "asr_only":
        return sa
    if strategy == "caption_only":
        return sb
    table = np.asarray(LEGACY_CLASS_WEIGHTS[strategy], dtype=np.float32)
    cls = np.asarray(row_class, dtype=np.int64) & 3
    return (table[cls, 0] * sa + table[cls, 1] * sb).astype(np.float32)


def class_weight_scores(q, asr_rows, caption_rows, has_asr, has_caption, row_class, class_weights) -> np.ndarray:
    """The general form the engine computes (include/cab.h, cab_score_all): fp32
    `w[c][0] * s_asr + w[c][1] * s_caption` with each operation rounded separately."""
    sa = cosine_rows(q, asr_rows, np.asarray(has_asr, dtype=bool))
    sb = cosine_rows(q, caption_rows, np.asarray(has_caption, dtype=bool))
    table = np.asarray(class_weights, dtype=np.float32).reshape(4, 2)
    cls = np.asarray(row_class, dtype=np.int64) & 3
    return (table[cls, 0] * sa + table[cls, 1] * sb).astype(np.float32)


# ---- clean_audio_search.py `search_audio` (:293-320) ----------------------------------------------
def clean_library(seed: int, n: int, n_queries: int, plants: int):
    """Unit-length synthetic embeddings for the clean_audio_search.py tests: (asr, caption, combined,
    has_asr, has_caption, queries), all float32, regenerated from the seed."""
    from multimodal_audio_search_b200 import synth
    a, c, f, _ = synth.library(seed, n, n_queries, plants, True)
    m, _, _, _ = synth.library(seed + 1000, n, n_queries, plants, False)
    q = synth.raw_queries(seed, 0, n_queries)
    qm = synth.raw_queries(seed + 1000, 0, n_queries)

    def unit(x):
        x = np.asarray(x, dtype=np.float32)
        nrm = np.sqrt(np.einsum("ij,ij->i", x, x))
        nrm[nrm == 0] = 1
        return (x / nrm[:, None]).astype(np.float32)
    # the combined corpus is planted for the second library's queries: blend both query sets so
    # every mode has neighbours for query i
    return unit(a), unit(c), unit(m), (f & 1).astype(bool), (f & 2).astype(bool), unit(q), unit(qm)


def clean_search(q, rows, has, k: int = 10, threshold: float = 0.1):
    """Raw dot product (no normalisation, :306-310) of the query with one stored embedding per
    segment, 0.0 where it is missing, strict `> 0.1` (:312), stable descending sort (:319), top 10
    (:320).  Returns (indices, similarities as Python-float-valued float64)."""
    q = np.asarray(q, dtype=np.float32)
    rows = np.asarray(rows, dtype=np.float32)
    sims = np.array([float(np.dot(q, rows[i])) if has[i] else 0.0 for i in range(len(rows))], dtype=np.float64)
    passing = np.nonzero(sims > threshold)[0]
    order = passing[np.argsort(-sims[passing], kind="stable")][:k]
    return order.astype(np.int64), sims[order]
