"""Drop-in for the search path of the reference's `DualPipelineAudioSearch`
(/root/reference/audio_search.py:87-699).

Two ways to use it:

* `accelerate(search_system)` -- patch an existing reference object (the one Streamlit keeps in
  `st.session_state.search_system`, :708-711): only `search_with_fusion` is replaced; the
  reference's own `_analyze_query_for_weights`, `text_embedder`, `stats`, `audio_segments`
  keep being used, so Whisper / MiniLM / the UI stay the reference's code.
* `DualPipelineAudioSearch` -- a standalone object with the same search-side attributes
  (`audio_segments`, `text_embedder`, `stats['search_pipeline']`, `search_with_fusion`,
  `_analyze_query_for_weights`) for callers that do not have the reference installed.

`search_with_fusion(query)` returns what the reference returns: `(results[:10], weight_info)`,
each result being `{**segment, asr_similarity, audio_similarity, fusion_score,
effective_asr_weight, effective_audio_weight, query_asr_weight, query_audio_weight}` (:673-682)
with Python floats; `([], {})` on an empty library before any stats update (:626-627).
The per-segment loop, threshold, sort and slice (:639-685, :699) run on the GPU (libcab.so).
"""
from __future__ import annotations

import time
import types
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

from . import query_weights
from .index import SegmentIndex
from .segment_table import SegmentRecord, SegmentTable

TOP_K = 10          # audio_search.py:699
THRESHOLD = 0.1     # audio_search.py:672
DIM = 384


@dataclass
class PipelineStats:
    """Same counters and update rule as the reference's PipelineStats (audio_search.py:23-48)."""
    pipeline_name: str
    model_name: str
    total_calls: int = 0
    total_processing_time: float = 0.0
    avg_processing_time: float = 0.0
    success_rate: float = 1.0
    successful_extractions: int = 0
    failed_extractions: int = 0

    def update(self, processing_time: float, success: bool):
        self.total_calls += 1
        self.total_processing_time += processing_time
        self.avg_processing_time = self.total_processing_time / self.total_calls
        if success:
            self.successful_extractions += 1
        else:
            self.failed_extractions += 1
        self.success_rate = self.successful_extractions / self.total_calls


def _embedding_row(e) -> np.ndarray:
    a = np.asarray(e, dtype=np.float32).reshape(-1)
    if a.shape[0] != DIM:
        raise ValueError(f"Incompatible dimension for X and Y matrices: X.shape[1] == {DIM} "
                         f"while Y.shape[1] == {a.shape[0]}")
    return a


class _DeviceLibrary:
    """Keeps a SegmentIndex in step with an append-only `audio_segments` list (:797)."""

    def __init__(self, dtype: str = "fp32", device: int = 0, tensor_core_batches: bool = False):
        self.dtype, self.device = dtype, device
        # fp32 library: keep bf16 shadows so that batches (search_many, SearchBatcher, search_batch)
        # are preselected on the tensor cores and re-scored exactly (SegmentIndex.enable_tensor_core_batches)
        self.tensor_core_batches = tensor_core_batches and dtype == "fp32"
        self.index: SegmentIndex | None = None
        self.n_synced = 0
        self._last_seg = None
        self._table_generation = -1

    def _sync_table(self, table: SegmentTable) -> SegmentIndex:
        """Columnar library: the table hands over the embeddings of its new rows once and
        forgets them; rows loaded from a file are in the index already (load_library)."""
        fresh = table is not self._last_seg or table.generation != self._table_generation
        if self.index is None:
            self.index = self._new_index(len(table))
        elif fresh:
            self.index.clear()
        if fresh:
            self._last_seg, self._table_generation = table, table.generation
            table.bind_index(self.index)
        if table.n_pending:
            row0, asr, audio, flags = table.pending_arrays()
            if row0 != len(self.index):
                raise RuntimeError(f"segment table and device index are out of step "
                                   f"(table rows before the new ones: {row0}, index rows: {len(self.index)})")
            self.index.append(asr, audio, flags)                      # ValueError on NaN/Inf: nothing appended,
            table.commit_pending(asr.shape[0])                        # nothing forgotten -- the next search retries
        if len(self.index) != len(table):
            raise RuntimeError(f"segment table has {len(table)} rows, the device index {len(self.index)}")
        self.n_synced = len(table)
        return self.index

    def _new_index(self, n: int) -> SegmentIndex:
        index = SegmentIndex(self.dtype, capacity=max(1024, n), device=self.device)
        if self.tensor_core_batches:
            index.enable_tensor_core_batches()
        return index

    def adopt(self, index: SegmentIndex, table: SegmentTable) -> None:
        """Take a loaded (index, table) pair as the synced state."""
        if len(index) != len(table):
            raise ValueError(f"index file has {len(index)} rows, segment table {len(table)}")
        if self.index is not None:
            self.index.close()
        self.index, self.dtype = index, index.dtype
        if self.tensor_core_batches and index.dtype == "fp32":
            index.enable_tensor_core_batches()
        self._last_seg, self._table_generation = table, table.generation
        table.bind_index(index)
        self.n_synced = len(table)

    def sync(self, segments) -> SegmentIndex:
        if isinstance(segments, SegmentTable):
            return self._sync_table(segments)
        if self.index is None:
            self.index = self._new_index(len(segments))
        # the reference only ever appends; if the list was replaced or shrunk, rebuild
        if self.n_synced > len(segments) or (self.n_synced and segments[self.n_synced - 1] is not self._last_seg):
            self.index.clear()
            self.n_synced = 0
        n_new = len(segments) - self.n_synced
        if n_new > 0:
            new = segments[self.n_synced:]
            asr = np.zeros((n_new, DIM), dtype=np.float32)
            audio = np.zeros((n_new, DIM), dtype=np.float32)
            flags = np.zeros(n_new, dtype=np.uint8)
            for i, seg in enumerate(new):
                if seg["asr_embedding"] is not None:                  # :644
                    asr[i] = _embedding_row(seg["asr_embedding"])
                if seg["audio_embedding"] is not None:                # :649
                    audio[i] = _embedding_row(seg["audio_embedding"])
                flags[i] = (1 if seg["asr_success"] else 0) | (2 if seg["audio_success"] else 0)
            self.index.append(asr, audio, flags)                      # ValueError on NaN/Inf
            self.n_synced = len(segments)
            self._last_seg = segments[-1]
        return self.index


def _materialise(self, res, qi: int, asr_weight: float, audio_weight: float) -> List[Dict]:
    """Result dicts of query row `qi` of a SearchResult, as the reference builds them (:653-682)."""
    results = []
    for j in range(int(res.count[qi])):
        segment = self.audio_segments[int(res.indices[qi, j])]
        asr_similarity = float(res.asr_sim[qi, j])                    # :646
        audio_similarity = float(res.audio_sim[qi, j])                # :651
        effective_asr_weight = asr_weight if segment["asr_success"] else 0       # :656-657
        effective_audio_weight = audio_weight if segment["audio_success"] else 0
        total_weight = effective_asr_weight + effective_audio_weight
        effective_asr_weight /= total_weight                          # :663-664
        effective_audio_weight /= total_weight
        fusion_score = (effective_asr_weight * asr_similarity +
                        effective_audio_weight * audio_similarity)    # :667-670
        scores = {
            "asr_similarity": asr_similarity,
            "audio_similarity": audio_similarity,
            "fusion_score": fusion_score,
            "effective_asr_weight": effective_asr_weight,
            "effective_audio_weight": effective_audio_weight,
            "query_asr_weight": asr_weight,
            "query_audio_weight": audio_weight,
        }
        # :673-682; a columnar record keeps audio_data / embeddings lazy until the UI reads them
        results.append(segment.with_fields(**scores) if isinstance(segment, SegmentRecord)
                       else {**segment, **scores})
    return results


def _b200_search_with_fusion(self, query: str) -> Tuple[List[Dict], Dict]:
    """GPU-backed body of search_with_fusion; `self` is a reference-compatible engine."""
    if not self.audio_segments:                                       # :626-627
        return [], {}
    start_time = time.time()
    asr_weight, audio_weight, weight_analysis = self._analyze_query_for_weights(query)   # :632
    query_embedding = _embedding_row(self.text_embedder.encode(query))                   # :635
    batcher = getattr(self, "_cab_batcher", None)
    if batcher is not None:           # shared library: this session's query rides in a batch with the others'
        if self._cab_library.n_synced != len(self.audio_segments):    # new segments: append between two batches
            batcher.run_exclusive(lambda: self._cab_library.sync(self.audio_segments))
        res = batcher.search(query_embedding, asr_weight, audio_weight, k=TOP_K, threshold=THRESHOLD)
    else:
        index = self._cab_library.sync(self.audio_segments)
        res = index.search(query_embedding[None, :], asr_weight, audio_weight, k=TOP_K, threshold=THRESHOLD)
    results = _materialise(self, res, 0, asr_weight, audio_weight)
    processing_time = time.time() - start_time
    self.stats["search_pipeline"].update(processing_time, success=len(results) > 0)      # :688-689
    weight_info = {"asr_weight": asr_weight, "audio_weight": audio_weight,
                   "analysis": weight_analysis, "query": query}       # :692-697
    return results, weight_info


def _b200_search_many(self, queries: List[str]) -> List[Tuple[List[Dict], Dict]]:
    """`search_with_fusion` for several query strings with ONE scan batch (one corpus pass per 4
    queries; the tensor-core scan from 4 queries on a bf16 or shadowed fp32 library): the list of what
    `search_with_fusion(q)` returns for each q, stats updated once per query."""
    if not self.audio_segments:
        return [([], {}) for _ in queries]
    if not queries:
        return []
    start_time = time.time()
    analysed = [self._analyze_query_for_weights(q) for q in queries]                     # :632
    vectors = np.stack([_embedding_row(v) for v in self.text_embedder.encode(list(queries))])   # :635, one forward
    index = self._cab_library.sync(self.audio_segments)
    res = index.search(vectors, [a[0] for a in analysed], [a[1] for a in analysed], k=TOP_K, threshold=THRESHOLD)
    out = []
    for qi, (query, (asr_weight, audio_weight, weight_analysis)) in enumerate(zip(queries, analysed)):
        results = _materialise(self, res, qi, asr_weight, audio_weight)
        out.append((results, {"asr_weight": asr_weight, "audio_weight": audio_weight,
                              "analysis": weight_analysis, "query": query}))
    share = (time.time() - start_time) / len(queries)
    for results, _ in out:
        self.stats["search_pipeline"].update(share, success=len(results) > 0)
    return out


def accelerate(search_system, dtype: str = "fp32", device: int = 0, columnar: bool = False,
               tensor_core_batches: bool = False):
    """Replace `search_system.search_with_fusion` (a reference DualPipelineAudioSearch, or any
    object with the same attributes) by the B200 path.  Returns the same object.

    `columnar=True` also replaces the `audio_segments` list by a SegmentTable holding the same
    records (the app's `audio_segments.extend(segments)`, :797, keeps working): embeddings then
    live only in HBM and the library can be saved with `save_library`.
    `tensor_core_batches=True` (fp32): batched calls (`search_many`, `enable_batching`,
    `search_batch`) are preselected on the tensor cores from bf16 shadow rows and re-scored exactly."""
    search_system._cab_library = _DeviceLibrary(dtype, device, tensor_core_batches)
    if columnar and not isinstance(search_system.audio_segments, SegmentTable):
        search_system.audio_segments = SegmentTable.from_segments(search_system.audio_segments)
    search_system.search_with_fusion = types.MethodType(_b200_search_with_fusion, search_system)
    search_system.search_many = types.MethodType(_b200_search_many, search_system)
    return search_system


def save_library(search_system, path: str, audio: bool = True) -> None:
    """Persist an accelerated engine's library: `path` (embeddings + flags as resident in HBM,
    SegmentIndex.save), `path + '.meta'` (columnar records) and `path + '.meta.audio'` (samples).
    The reference has no counterpart: its library dies with the Streamlit session (:708-711)."""
    if not search_system.audio_segments:
        raise ValueError("the library is empty")
    index = search_system._cab_library.sync(search_system.audio_segments)
    segments = search_system.audio_segments
    table = segments if isinstance(segments, SegmentTable) else SegmentTable.from_segments(segments)
    index.save(path)
    table.save(path + ".meta", audio=audio)


def load_library(search_system, path: str, device: int | None = None) -> None:
    """Inverse of save_library: `audio_segments` becomes a memory-mapped SegmentTable, the
    embeddings go straight from the file to HBM; no per-segment Python work."""
    lib = search_system._cab_library
    table = SegmentTable.load(path + ".meta")
    index = SegmentIndex.load(path, device=lib.device if device is None else device)
    lib.adopt(index, table)
    search_system.audio_segments = table


class DualPipelineAudioSearch:
    """Search-side mirror of the reference class (:87-140, :457-699).  Model loading and audio
    ingest are the reference's business: populate `audio_segments` with its segment records
    (:275-294) and set `text_embedder` to anything with `.encode(str) -> float32[384]`."""

    def __init__(self, dtype: str = "fp32", device: int = 0, text_embedder=None, tensor_core_batches: bool = False):
        self.text_embedder = text_embedder
        self.stats = {"search_pipeline": PipelineStats("Search Pipeline", "Cosine Similarity")}   # :107
        self.audio_segments: List[Dict] = []                                                      # :115
        self._cab_library = _DeviceLibrary(dtype, device, tensor_core_batches)

    def _analyze_query_for_weights(self, query: str):
        return query_weights.analyze_query_for_weights(query)

    search_with_fusion = _b200_search_with_fusion
    search_many = _b200_search_many

    def enable_batching(self, max_batch: int = 256, max_wait_s: float = 0.0):
        """Concurrent `search_with_fusion` calls (threads sharing this engine) are coalesced into
        batched GPU calls by a SearchBatcher; returns it (call `.close()` to stop)."""
        from .batcher import SearchBatcher
        index = self._cab_library.sync(self.audio_segments)
        self._cab_batcher = SearchBatcher(index, max_batch=max_batch, max_wait_s=max_wait_s)
        return self._cab_batcher

    # -- beyond the reference: a library that outlives the session ---
    def save_library(self, path: str, audio: bool = True) -> None:
        save_library(self, path, audio=audio)

    def load_library(self, path: str) -> None:
        load_library(self, path)

    # -- beyond the reference: many queries per call (BASELINE configs 3-5) ----------------------
    def search_batch(self, query_vectors, asr_weights, audio_weights, k: int = TOP_K,
                     threshold: float = THRESHOLD, path: str = "auto"):
        """Top-k for a batch of already-embedded queries; returns a SearchResult."""
        index = self._cab_library.sync(self.audio_segments)
        return index.search(query_vectors, asr_weights, audio_weights, k=k, threshold=threshold, path=path)
