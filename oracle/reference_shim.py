"""TEST INFRASTRUCTURE -- never imported by the product package.

Loads the UNMODIFIED reference `/root/reference/audio_search.py` in this build container so its
own `search_with_fusion` (audio_search.py:624-699) and `_analyze_query_for_weights` (:457-622)
can be executed on chosen inputs.  Used only to (1) pin the numpy restatement in
`oracle/numpy_oracle.py` and (2) mint the golden fixtures under `tests/golden/`
(`oracle/make_golden.py`).  `/root/reference` does not exist on the GPU box; nothing marked
`gpu`, `smoke()` or `bench.py` may call this.

The reference imports streamlit / librosa / sentence_transformers at module top
(audio_search.py:7,9,12).  None is installed and none is touched by the two functions we run, so
they are replaced by empty stub modules.  `transformers` must be imported BEFORE the stubs exist
(it probes `librosa.__spec__`), see SURVEY.md Appendix A.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_FILE = os.environ.get("CAB_REFERENCE_FILE", "/root/reference/audio_search.py")


def available() -> bool:
    return os.path.exists(REFERENCE_FILE)


_module = None


def load():
    """Import the reference module once (stubbing the three absent third-party packages)."""
    global _module
    if _module is not None:
        return _module
    if not available():
        raise FileNotFoundError(REFERENCE_FILE)
    from transformers import pipeline, WhisperProcessor, WhisperForConditionalGeneration  # noqa: F401
    for name in ("streamlit", "librosa", "sentence_transformers"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["sentence_transformers"], "SentenceTransformer"):
        sys.modules["sentence_transformers"].SentenceTransformer = object
    spec = importlib.util.spec_from_file_location("reference_audio_search", REFERENCE_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _module = mod
    return mod


LEGACY_FILE = os.path.join(os.path.dirname(REFERENCE_FILE), "previous_iterations", "streamlit_app.py")
_legacy = None


def legacy_available() -> bool:
    return os.path.exists(LEGACY_FILE)


def load_legacy():
    """Import the reference's earlier engine (previous_iterations/streamlit_app.py) for its
    `UnifiedAudioSearch.search` (:173-223).  Its class body applies `@st.cache_resource`, so the
    streamlit stub gets a pass-through decorator; `main()` is guarded by `__name__`."""
    global _legacy
    if _legacy is not None:
        return _legacy
    if not legacy_available():
        raise FileNotFoundError(LEGACY_FILE)
    load()                                          # installs the stub modules
    st = sys.modules["streamlit"]
    if not hasattr(st, "cache_resource"):
        st.cache_resource = lambda f=None, **_kw: f if f is not None else (lambda g: g)
    spec = importlib.util.spec_from_file_location("reference_legacy_app", LEGACY_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _legacy = mod
    return mod


class ListEmbedder:
    """The earlier engine calls `sentence_model.encode([text])[0]` (:184)."""

    def __init__(self, table: dict[str, np.ndarray]):
        self.table = {k: np.asarray(v, dtype=np.float32) for k, v in table.items()}

    def encode(self, texts, **_kw):
        return np.stack([self.table[t] for t in texts])


def legacy_database(asr, caption, has_asr, has_caption, good_speech) -> list[dict]:
    """The earlier engine's database items (streamlit_app.py:322-336): embeddings or None, and a
    transcript whose stripped length decides the adaptive weights (:216)."""
    db = []
    for i in range(len(has_asr)):
        db.append({
            "filename": "synthetic.wav", "chunk_id": i, "start_time": 10.0 * i, "end_time": 10.0 * i + 10.0,
            "asr_embedding": np.asarray(asr[i], dtype=np.float32) if has_asr[i] else None,
            "caption_embedding": np.asarray(caption[i], dtype=np.float32) if has_caption[i] else None,
            "asr_transcription": ("a clearly spoken sentence %d" % i) if good_speech[i] else " uh ",
            "audio_caption": "some sound",
            "has_asr": bool(has_asr[i]), "has_caption": bool(has_caption[i]),
        })
    return db


def legacy_search(query_text: str, query_vec, database, strategy: str) -> np.ndarray:
    """Run the reference's own `UnifiedAudioSearch.search` on `database`."""
    mod = load_legacy()
    eng = mod.UnifiedAudioSearch()
    eng.sentence_model = ListEmbedder({query_text: query_vec})
    return eng.search(query_text, database, strategy)


CLEAN_FILE = os.path.join(os.path.dirname(REFERENCE_FILE), "previous_iterations", "clean_audio_search.py")
_clean = None


def clean_available() -> bool:
    return os.path.exists(CLEAN_FILE)


def load_clean():
    """Import previous_iterations/clean_audio_search.py for its `search_audio` (:293-320)."""
    global _clean
    if _clean is not None:
        return _clean
    if not clean_available():
        raise FileNotFoundError(CLEAN_FILE)
    load()                                          # installs the stub modules
    spec = importlib.util.spec_from_file_location("reference_clean_app", CLEAN_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _clean = mod
    return mod


def clean_database(asr, caption, combined, has_asr, has_caption) -> list[dict]:
    """Segments as clean_audio_search.py stores them (:175-187); a segment always has a combined
    embedding (it is only stored when some text exists, :171)."""
    db = []
    for i in range(len(has_asr)):
        db.append({
            "segment_id": f"seg_{i}", "start_time": 10.0 * i, "end_time": 10.0 * i + 10.0,
            "asr_text": f"words {i}" if has_asr[i] else "", "caption_text": f"sound {i}" if has_caption[i] else "",
            "combined_text": f"words sound {i}",
            "asr_embedding": np.asarray(asr[i], dtype=np.float32) if has_asr[i] else None,
            "caption_embedding": np.asarray(caption[i], dtype=np.float32) if has_caption[i] else None,
            "combined_embedding": np.asarray(combined[i], dtype=np.float32),
            "audio_data": None, "sample_rate": 16000,
        })
    return db


def clean_search(query_text: str, query_vec, database, search_mode: str):
    """Run the reference's own `UnifiedAudioSearch.search_audio` (clean_audio_search.py)."""
    mod = load_clean()
    eng = mod.UnifiedAudioSearch.__new__(mod.UnifiedAudioSearch)
    eng.text_embedder = FakeEmbedder({query_text: query_vec})
    eng.audio_database = database
    return eng.search_audio(query_text, search_mode)


class FakeEmbedder:
    """Stands in for SentenceTransformer: `.encode(text)` returns the vector registered for it."""

    def __init__(self, table: dict[str, np.ndarray]):
        self.table = {k: np.asarray(v, dtype=np.float32) for k, v in table.items()}

    def encode(self, text, **_kw):
        if isinstance(text, (list, tuple)):          # SentenceTransformer.encode(list) -> [n, 384]
            return np.stack([self.table[t] for t in text])
        return self.table[text].copy()


def make_segment(i: int, asr_emb, audio_emb, asr_success=None, audio_success=None) -> dict:
    """A segment record with the reference's 12 keys (audio_search.py:275-294)."""
    asr_success = (asr_emb is not None) if asr_success is None else asr_success
    audio_success = (audio_emb is not None) if audio_success is None else audio_success
    return {
        "segment_id": f"seg_{i}",
        "start_time": 10.0 * i,
        "end_time": 10.0 * i + 10.0,
        "duration": 10.0,
        "asr_text": f"asr text {i}" if asr_success else "",
        "asr_embedding": None if asr_emb is None else np.asarray(asr_emb, dtype=np.float32),
        "asr_success": bool(asr_success),
        "audio_description": f"audio description {i}" if audio_success else "",
        "audio_embedding": None if audio_emb is None else np.asarray(audio_emb, dtype=np.float32),
        "audio_success": bool(audio_success),
        "audio_data": None,
        "sample_rate": 16000,
    }


def segments_from_arrays(asr, audio, flags, has_asr=None, has_audio=None) -> list[dict]:
    """Build the reference's list-of-dict library from row matrices + success flags.
    By default an embedding is present iff its pipeline succeeded (audio_search.py:282-289)."""
    n = len(flags)
    segs = []
    for i in range(n):
        sa, sb = bool(flags[i] & 1), bool(flags[i] & 2)
        ha = sa if has_asr is None else bool(has_asr[i])
        hb = sb if has_audio is None else bool(has_audio[i])
        segs.append(make_segment(i, asr[i] if ha else None, audio[i] if hb else None, sa, sb))
    return segs


def reference_engine(segments: list[dict], queries: dict[str, np.ndarray]):
    """A reference `DualPipelineAudioSearch` holding `segments`, with query text -> vector."""
    ref = load()
    eng = ref.DualPipelineAudioSearch()
    eng.text_embedder = FakeEmbedder(queries)
    eng.audio_segments = segments
    return eng


def reference_weights(query: str):
    ref = load()
    eng = ref.DualPipelineAudioSearch.__new__(ref.DualPipelineAudioSearch)
    return eng._analyze_query_for_weights(query)
