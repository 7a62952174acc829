"""Drop-in for the search method of the reference's EARLIER engine, `UnifiedAudioSearch.search`
(/root/reference/previous_iterations/streamlit_app.py:173-223) -- SURVEY.md section 8(f) rank 4.

That version scores every database item and returns the whole vector, `np.array(similarities)`
(:223); its UI takes `np.argsort(similarities)[::-1][:top_k]` (:410).  Three strategies:

    asr_only      cos(query, asr_embedding)                                   (:188-193)
    caption_only  cos(query, caption_embedding)                               (:195-200)
    adaptive      0.7 cos_asr + 0.3 cos_caption  if len(transcript.strip()) > 10
                  0.2 cos_asr + 0.8 cos_caption  otherwise  (any other strategy string)  (:202-219)

A missing embedding (`None`) counts 0.0.  Here the database's embeddings live in a SegmentIndex
(flag bit0/bit1 = embedding present, bits 2-3 = weight class: 1 for a transcript longer than 10
characters) and one `cab_score_all` launch produces the N scores on the GPU.

* `accelerate_legacy(system)` patches a reference `UnifiedAudioSearch` object in place;
* `UnifiedAudioSearch` is a standalone mirror of its search-side surface (`sentence_model`,
  `search(query_text, audio_database, strategy)`).
"""
from __future__ import annotations

import types
from typing import Dict, List

import numpy as np

from .index import SegmentIndex

DIM = 384
GOOD_SPEECH_CLASS = 1
# {w_asr, w_caption} per row weight class
CLASS_WEIGHTS = {
    "asr_only": ((1.0, 0.0),) * 4,
    "caption_only": ((0.0, 1.0),) * 4,
    "adaptive": ((0.2, 0.8), (0.7, 0.3), (0.2, 0.8), (0.2, 0.8)),             # :219 / :217
}


def item_flags(item: Dict) -> int:
    """Flag byte of a database item (:322-336)."""
    f = (1 if item.get("asr_embedding") is not None else 0) | (2 if item.get("caption_embedding") is not None else 0)
    if len(item.get("asr_transcription", "").strip()) > 10:                   # :215-216
        f |= GOOD_SPEECH_CLASS << 2
    return f


def _row(e) -> np.ndarray:
    a = np.asarray(e, dtype=np.float32).reshape(-1)
    if a.shape[0] != DIM:
        raise ValueError(f"Incompatible dimension for X and Y matrices: X.shape[1] == {DIM} "
                         f"while Y.shape[1] == {a.shape[0]}")
    return a


class _DeviceDatabase:
    """Keeps a SegmentIndex in step with an append-only `audio_database` list (:338)."""

    def __init__(self, dtype: str = "fp32", device: int = 0):
        self.dtype, self.device = dtype, device
        self.index: SegmentIndex | None = None
        self.n_synced = 0
        self._list = None
        self._last = None

    def sync(self, database: List[Dict]) -> SegmentIndex:
        if self.index is None:
            self.index = SegmentIndex(self.dtype, capacity=max(1024, len(database)), device=self.device)
        replaced = database is not self._list or self.n_synced > len(database) or \
            (self.n_synced and database[self.n_synced - 1] is not self._last)
        if replaced:
            self.index.clear()
            self.n_synced = 0
            self._list = database
        n_new = len(database) - self.n_synced
        if n_new > 0:
            asr = np.zeros((n_new, DIM), dtype=np.float32)
            cap = np.zeros((n_new, DIM), dtype=np.float32)
            flags = np.zeros(n_new, dtype=np.uint8)
            for i, item in enumerate(database[self.n_synced:]):
                if item.get("asr_embedding") is not None:
                    asr[i] = _row(item["asr_embedding"])
                if item.get("caption_embedding") is not None:
                    cap[i] = _row(item["caption_embedding"])
                flags[i] = item_flags(item)
            self.index.append(asr, cap, flags)                                # ValueError on NaN/Inf
            self.n_synced = len(database)
            self._last = database[-1]
        return self.index


def _b200_search(self, query_text, audio_database, strategy="adaptive"):
    """GPU-backed body of `UnifiedAudioSearch.search`: float64 [N] similarities."""
    query_embedding = _row(self.sentence_model.encode([query_text])[0])       # :184
    if len(audio_database) == 0:
        return np.array([])                                                   # np.array([]) of :223
    index = self._cab_database.sync(audio_database)
    weights = CLASS_WEIGHTS.get(strategy, CLASS_WEIGHTS["adaptive"])          # the `else:` of :202
    scores = index.score_all(query_embedding[None, :], weights, out=index.pinned_scores(1))
    return scores[0].astype(np.float64)       # an owned float64 vector, like np.array(similarities)


def accelerate_legacy(system, dtype: str = "fp32", device: int = 0):
    """Replace `system.search` (a reference `UnifiedAudioSearch`, or any object with a
    `sentence_model`) by the B200 path.  Returns the same object."""
    system._cab_database = _DeviceDatabase(dtype, device)
    system.search = types.MethodType(_b200_search, system)
    return system


class UnifiedAudioSearch:
    """Search-side mirror of the earlier reference class (:49-70, :173-223).  Model loading and
    audio chunking stay the reference's business: pass `sentence_model` (anything with
    `.encode([text]) -> [float32[384]]`) and database items shaped like :322-336."""

    def __init__(self, dtype: str = "fp32", device: int = 0, sentence_model=None):
        self.sentence_model = sentence_model
        self._cab_database = _DeviceDatabase(dtype, device)

    search = _b200_search

    @staticmethod
    def top_indices(similarities: np.ndarray, top_k: int) -> np.ndarray:
        """What the reference UI does with the vector (:410)."""
        return np.argsort(similarities)[::-1][:top_k]
