"""Keyword-derived fusion weights for a query -- host-side twin of the reference's
`DualPipelineAudioSearch._analyze_query_for_weights` (/root/reference/audio_search.py:457-622).

The rule is substring (not token) counting of two keyword lists in the lower-cased query, then a
four-branch weight formula clamped to [0.2, 0.8].  It costs tens of microseconds and its outputs
must be bit-identical Python floats (e.g. 1.0 - 0.7 == 0.30000000000000004), so it stays on the
host in float64; the device only sees the two resulting weights per query.
"""
from __future__ import annotations

from typing import Tuple

from .keyword_table import ASR_KEYWORDS, AUDIO_KEYWORDS


def _bucket(table):
    """Keywords by their first two characters: a keyword can only occur in a query that contains
    that bigram, so a query of n characters tests the few keywords behind its <= n-1 bigrams instead
    of all 586 (one-character keywords, if any, are tested always)."""
    by_bigram, short = {}, []
    for kw, n in table.items():
        if len(kw) >= 2:
            by_bigram.setdefault(kw[:2], []).append((kw, n))
        else:
            short.append((kw, n))
    return by_bigram, short


_ASR_BUCKETS, _AUDIO_BUCKETS = _bucket(ASR_KEYWORDS), _bucket(AUDIO_KEYWORDS)


def _count(query_lower: str, buckets) -> int:
    by_bigram, short = buckets
    total = sum(n for kw, n in short if kw in query_lower)
    for bg in {query_lower[i:i + 2] for i in range(len(query_lower) - 1)}:
        for kw, n in by_bigram.get(bg, ()):
            if kw in query_lower:
                total += n
    return total


def count_matches(query_lower: str) -> Tuple[int, int]:
    """(asr_matches, audio_matches): every listed occurrence of a keyword that is a substring of
    the query counts (:586-587) -- `sum(n for kw, n in table.items() if kw in query)`, evaluated
    through the bigram buckets."""
    return _count(query_lower, _ASR_BUCKETS), _count(query_lower, _AUDIO_BUCKETS)


def count_matches_plain(query_lower: str) -> Tuple[int, int]:
    """The literal form (every keyword tested); kept as the cross-check of the bucketed one."""
    asr = sum(n for kw, n in ASR_KEYWORDS.items() if kw in query_lower)
    audio = sum(n for kw, n in AUDIO_KEYWORDS.items() if kw in query_lower)
    return asr, audio


def weights_from_matches(asr_matches: int, audio_matches: int) -> Tuple[float, float, str]:
    if asr_matches == 0 and audio_matches == 0:                       # :593-596
        return 0.5, 0.5, "Balanced (no specific keywords detected)"
    if audio_matches == 0:                                            # :598-603
        asr_weight = 0.5 + (0.3 * min(asr_matches / 3.0, 1.0))
        return asr_weight, 1.0 - asr_weight, f"ASR-focused ({asr_matches} speech keywords)"
    if asr_matches == 0:                                              # :605-610
        audio_weight = 0.5 + (0.3 * min(audio_matches / 3.0, 1.0))
        return 1.0 - audio_weight, audio_weight, f"Audio-focused ({audio_matches} audio keywords)"
    asr_ratio = asr_matches / (asr_matches + audio_matches)           # :612-620
    asr_weight = 0.2 + (asr_ratio * 0.6)
    return asr_weight, 1.0 - asr_weight, f"Mixed query (ASR:{asr_matches}, Audio:{audio_matches})"


def analyze_query_for_weights(query: str) -> Tuple[float, float, str]:
    """(asr_weight, audio_weight, analysis) exactly as the reference returns them (:622)."""
    return weights_from_matches(*count_matches(query.lower()))
