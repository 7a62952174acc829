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

The other earlier engine, `previous_iterations/clean_audio_search.py`, searches ONE stored embedding
per segment (`search_audio(query, search_mode)`, :293-320: "combined" / "asr" / "caption", raw dot
product, strict `> 0.1`, stable sort, top 10).  `accelerate_clean(system)` / `CleanAudioSearch`
serve it with the top-k scan: a single-corpus search is the fused search with weights (1, 0) or
(0, 1) (rows lacking that embedding are skipped, audio_search.py:659-661 rule), on unit-length
embeddings the raw dot product is the cosine, and the few winners are re-scored on the host with
the reference's own expression so the returned similarities are the reference's floats.
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


# ---- clean_audio_search.py: search_audio(query, search_mode) ------------------------------------------
CLEAN_TOP_K = 10            # clean_audio_search.py:320
CLEAN_THRESHOLD = 0.1       # :312
_CLEAN_OVERFETCH = 64       # candidates taken from the GPU before the exact host re-score
_CLEAN_MARGIN = 0.01        # the over-fetch threshold sits this far below 0.1 (GPU fp32/bf16 vs the host re-score)
_CLEAN_KEYS = {"combined": "combined_embedding", "asr": "asr_embedding", "caption": "caption_embedding"}


class _CleanDatabase:
    """Two indexes in step with the append-only `audio_database` (:115-187): (asr, caption) rows
    in one, the combined embedding in the other (its second corpus stays empty)."""

    def __init__(self, dtype: str = "fp32", device: int = 0):
        self.dtype, self.device = dtype, device
        self.pair: SegmentIndex | None = None
        self.combined: SegmentIndex | None = None
        self.n_synced = 0
        self._list = None
        self._last = None

    def sync(self, database: List[Dict]):
        if self.pair is None:
            cap = max(1024, len(database))
            self.pair = SegmentIndex(self.dtype, capacity=cap, device=self.device)
            self.combined = SegmentIndex(self.dtype, capacity=cap, device=self.device)
            # :306 ranks by the RAW dot product: embeddings of any length are served (the index keeps
            # each row's original length; for unit-length MiniLM rows this is the cosine)
            self.pair.set_option("raw_dot", 1)
            self.combined.set_option("raw_dot", 1)
        replaced = database is not self._list or self.n_synced > len(database) or \
            (self.n_synced and database[self.n_synced - 1] is not self._last)
        if replaced:
            self.pair.clear()
            self.combined.clear()
            self.n_synced = 0
            self._list = database
        n_new = len(database) - self.n_synced
        if n_new > 0:
            rows = {key: np.zeros((n_new, DIM), dtype=np.float32) for key in _CLEAN_KEYS.values()}
            has = {key: np.zeros(n_new, dtype=np.uint8) for key in _CLEAN_KEYS.values()}
            for i, seg in enumerate(database[self.n_synced:]):
                for key in _CLEAN_KEYS.values():
                    if seg.get(key) is not None:
                        rows[key][i] = _row(seg[key])
                        has[key][i] = 1
            self.pair.append(rows["asr_embedding"], rows["caption_embedding"],
                             has["asr_embedding"] | (has["caption_embedding"] << 1))
            self.combined.append(rows["combined_embedding"], None, has["combined_embedding"])
            self.n_synced = len(database)
            self._last = database[-1]
        return self.pair, self.combined


def _b200_search_audio(self, query: str, search_mode: str = "combined") -> List[Dict]:
    """GPU-backed body of clean_audio_search.py `search_audio`."""
    if not self.audio_database:                                               # :295-296
        return []
    query_embedding = self.text_embedder.encode(query)                        # :299
    key = _CLEAN_KEYS.get(search_mode)
    if key is None:
        return []                                                             # every similarity stays 0.0 (:303-310)
    q = _row(query_embedding)
    pair, combined = self._cab_database.sync(self.audio_database)
    index, wa, wb = (combined, 1.0, 0.0) if search_mode == "combined" else \
        (pair, 1.0, 0.0) if search_mode == "asr" else (pair, 0.0, 1.0)
    res = index.search(q[None, :], wa, wb, k=_CLEAN_OVERFETCH, threshold=CLEAN_THRESHOLD - _CLEAN_MARGIN)
    scored = []
    for j in range(int(res.count[0])):
        i = int(res.indices[0, j])
        segment = self.audio_database[i]
        similarity = float(np.dot(query_embedding, segment[key]))             # :306-310, the reference's own float
        if similarity > CLEAN_THRESHOLD:                                      # :312
            scored.append((-similarity, i, segment, similarity))
    scored.sort(key=lambda t: (t[0], t[1]))                                   # :319 stable descending
    return [{**segment, "similarity": similarity} for _, _, segment, similarity in scored[:CLEAN_TOP_K]]   # :313-320


def accelerate_clean(system, dtype: str = "fp32", device: int = 0):
    """Replace `system.search_audio` (a clean_audio_search.py `UnifiedAudioSearch`, or any object
    with `text_embedder` and `audio_database`) by the B200 path.  Returns the same object."""
    system._cab_database = _CleanDatabase(dtype, device)
    system.search_audio = types.MethodType(_b200_search_audio, system)
    return system


class CleanAudioSearch:
    """Search-side mirror of clean_audio_search.py's `UnifiedAudioSearch` (:26-68, :293-320)."""

    def __init__(self, dtype: str = "fp32", device: int = 0, text_embedder=None):
        self.text_embedder = text_embedder
        self.audio_database: List[Dict] = []                                  # :66
        self._cab_database = _CleanDatabase(dtype, device)

    search_audio = _b200_search_audio
