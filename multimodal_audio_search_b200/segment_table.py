"""Columnar segment metadata: the host-side half of the library.

The reference keeps one Python dict per segment in `audio_segments` (audio_search.py:275-294) --
the embeddings, the texts, the times and the raw audio all live in that dict, and every search
result is `{**segment, ...}` (:673-682).  Here the embeddings live in HBM (SegmentIndex) and the
rest lives in columns: numeric columns as flat arrays, text columns as (offsets, utf-8 blob),
audio samples as one float32 blob.  A `SegmentTable`

* is a drop-in for the `audio_segments` list on the search path: `len()`, truthiness, indexing,
  iteration, `append(record)`, `extend(records)` (:797) -- a record is a reference segment dict;
* materialises a result in O(1) per row without touching the other N-1 rows; `audio_data` and
  the two embeddings are fetched lazily, on first access, by the `SegmentRecord` it returns
  (:873 reads `audio_data` only for the results the user expands);
* adds the `file` column the reference lacks (segment ids restart at `seg_0` for every uploaded
  file, :276, so a result cannot be traced back to its file there);
* saves to / loads from a sidecar file next to the index file (`SegmentIndex.save`), loading by
  memory-mapping: opening a 10 M-segment library costs no per-row work.

File layout (little endian): 64-byte header {magic "CABMETA1", u32 version, u32 reserved,
u64 n_rows, u64 directory_offset, u64 directory_bytes}, 64-byte-aligned column blocks, then a JSON
directory [{name, dtype, offset, count}].  Audio samples go to a second file, `<path>.audio`
(raw float32), addressed by the `audio_offsets` column.
"""
from __future__ import annotations

import json
import os
import struct
from collections.abc import Sequence
from typing import Dict, Iterable, List, Optional

import numpy as np

MAGIC = b"CABMETA1"
VERSION = 1
_HEADER = struct.Struct("<8sIIQQQ24x")       # 64 bytes
assert _HEADER.size == 64

DIM = 384

# reference record fields (audio_search.py:275-294) that are stored as columns
NUMERIC_COLUMNS = (("start_time", "<f8"), ("end_time", "<f8"), ("duration", "<f8"),
                   ("asr_success", "u1"), ("audio_success", "u1"), ("sample_rate", "<i4"))
TEXT_COLUMNS = ("segment_id", "asr_text", "audio_description", "file", "extra_json")
LAZY_FIELDS = ("asr_embedding", "audio_embedding", "audio_data")
# key order of a reference record, so a materialised record lists its keys the same way
RECORD_ORDER = ("segment_id", "start_time", "end_time", "duration", "asr_text", "asr_embedding",
                "asr_success", "audio_description", "audio_embedding", "audio_success",
                "audio_data", "sample_rate")
_KNOWN = set(RECORD_ORDER) | {"file"}


class SegmentRecord(dict):
    """One segment as the reference's dict, with `audio_data`, `asr_embedding` and
    `audio_embedding` fetched from the table on first access instead of being copied up front."""

    __slots__ = ("_table", "_row")

    def __init__(self, table: "SegmentTable", row: int, fields: Dict):
        super().__init__(fields)
        self._table, self._row = table, row

    # dict protocol, lazily completed --------------------------------------------------------
    def __missing__(self, key):
        if key in LAZY_FIELDS:
            value = self._table.fetch(self._row, key)
            dict.__setitem__(self, key, value)
            return value
        raise KeyError(key)

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in LAZY_FIELDS

    def _ordered_keys(self) -> List[str]:
        present = set(dict.keys(self)) | set(LAZY_FIELDS)
        head = [k for k in RECORD_ORDER if k in present]
        return head + [k for k in dict.keys(self) if k not in RECORD_ORDER]

    def keys(self):
        return self._ordered_keys()

    def __iter__(self):
        return iter(self._ordered_keys())

    def __len__(self):
        return len(self._ordered_keys())

    def values(self):
        return [self[k] for k in self._ordered_keys()]

    def items(self):
        return [(k, self[k]) for k in self._ordered_keys()]

    def copy(self) -> "SegmentRecord":
        return SegmentRecord(self._table, self._row, dict(dict.items(self)))

    def with_fields(self, **extra) -> "SegmentRecord":
        """`{**segment, **extra}` (:673-682) that keeps the lazy fields lazy."""
        r = self.copy()
        dict.update(r, extra)
        return r

    def materialize(self) -> Dict:
        """A plain dict with every field resolved (what `{**record}` gives, too)."""
        return {k: self[k] for k in self._ordered_keys()}

    def __eq__(self, other):
        return self.materialize() == (other.materialize() if isinstance(other, SegmentRecord) else other)

    __hash__ = None

    def __repr__(self):
        return f"SegmentRecord(row={self._row}, {dict.__repr__(self)})"


class _TextColumn:
    """Strings as (int64 offsets[n+1], utf-8 blob) for the loaded part + a Python list tail."""

    def __init__(self, offsets: Optional[np.ndarray] = None, blob: Optional[np.ndarray] = None):
        self.offsets = offsets if offsets is not None else np.zeros(1, dtype=np.int64)
        self.blob = blob if blob is not None else np.zeros(0, dtype=np.uint8)
        self.tail: List[str] = []

    @property
    def n_base(self) -> int:
        return len(self.offsets) - 1

    def __len__(self):
        return self.n_base + len(self.tail)

    def __getitem__(self, i: int) -> str:
        if i < self.n_base:
            lo, hi = int(self.offsets[i]), int(self.offsets[i + 1])
            return bytes(self.blob[lo:hi]).decode("utf-8") if hi > lo else ""
        return self.tail[i - self.n_base]

    def consolidated(self):
        if not self.tail:
            return np.asarray(self.offsets, dtype=np.int64), np.asarray(self.blob, dtype=np.uint8)
        enc = [s.encode("utf-8") for s in self.tail]
        lens = np.fromiter((len(b) for b in enc), dtype=np.int64, count=len(enc))
        offs = np.concatenate([np.asarray(self.offsets, dtype=np.int64),
                               int(self.offsets[-1]) + np.cumsum(lens)])
        blob = np.concatenate([np.asarray(self.blob, dtype=np.uint8),
                               np.frombuffer(b"".join(enc), dtype=np.uint8)])
        return offs, blob


class SegmentTable(Sequence):
    """Append-only columnar store of segment records; see the module docstring."""

    def __init__(self):
        self._num = {name: np.zeros(0, dtype=dt) for name, dt in NUMERIC_COLUMNS}      # loaded part
        self._num_tail: Dict[str, list] = {name: [] for name, _ in NUMERIC_COLUMNS}
        self._text = {name: _TextColumn() for name in TEXT_COLUMNS}
        self._audio_offsets = np.zeros(1, dtype=np.int64)       # loaded part, in samples
        self._audio_blob: Optional[np.ndarray] = None            # memmap of <path>.audio
        self._audio_tail: List[Optional[np.ndarray]] = []
        self._n_base = 0
        self._n = 0
        # embeddings of rows that are not on the device yet: (first_row, asr, audio) per row
        self._pending: List[tuple] = []
        self._pending_row0 = 0
        self._index = None                                       # SegmentIndex for embedding fetches
        self._fast = None                                        # memoryviews of the loaded columns (record())
        self.generation = 0                                      # bumped by clear()

    # ---- construction ------------------------------------------------------------------------
    @classmethod
    def from_segments(cls, segments: Iterable[Dict], file: Optional[str] = None) -> "SegmentTable":
        t = cls()
        t.extend(segments, file=file)
        return t

    @classmethod
    def from_columns(cls, n: int, *, start_time=None, end_time=None, asr_success=None,
                     audio_success=None, sample_rate=16000, segment_id=None, asr_text=None,
                     audio_description=None, file=None) -> "SegmentTable":
        """Bulk construction from arrays (synthetic / imported libraries): no per-row Python
        dicts.  The embeddings of such rows are expected to be in the index already."""
        t = cls()
        st = np.arange(n, dtype=np.float64) * 5.0 if start_time is None else np.asarray(start_time, dtype=np.float64)
        en = st + 10.0 if end_time is None else np.asarray(end_time, dtype=np.float64)
        t._num["start_time"], t._num["end_time"], t._num["duration"] = st, en, en - st
        t._num["asr_success"] = (np.ones(n, np.uint8) if asr_success is None else np.asarray(asr_success).astype(np.uint8))
        t._num["audio_success"] = (np.ones(n, np.uint8) if audio_success is None else np.asarray(audio_success).astype(np.uint8))
        t._num["sample_rate"] = np.broadcast_to(np.asarray(sample_rate, dtype=np.int32), (n,)).copy()
        for name, values in (("segment_id", segment_id), ("asr_text", asr_text),
                             ("audio_description", audio_description), ("file", file)):
            col = _TextColumn()
            if values is None and name == "segment_id":
                values = [f"seg_{i}" for i in range(n)] if n <= 100000 else None
            if values is None:
                col.offsets = np.zeros(n + 1, dtype=np.int64)
            else:
                col.tail = [str(v) for v in values]
                if len(col.tail) != n:
                    raise ValueError(f"column {name}: {len(col.tail)} values for {n} rows")
                col.offsets, col.blob = col.consolidated()
                col.tail = []
            t._text[name] = col
        t._text["extra_json"].offsets = np.zeros(n + 1, dtype=np.int64)
        t._audio_offsets = np.zeros(n + 1, dtype=np.int64)
        t._n_base = t._n = n
        t._pending_row0 = n
        for name, dt in NUMERIC_COLUMNS:
            if len(t._num[name]) != n:
                raise ValueError(f"column {name}: {len(t._num[name])} values for {n} rows")
        return t

    # ---- list protocol (what the reference does with `audio_segments`) ------------------------
    def __len__(self) -> int:
        return self._n

    def __bool__(self) -> bool:
        return self._n > 0

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self.record(j) for j in range(*i.indices(self._n))]
        i = int(i)
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError("segment index out of range")
        return self.record(i)

    def append(self, segment: Dict, file: Optional[str] = None) -> None:
        unknown = {k: v for k, v in segment.items() if k not in _KNOWN}
        tail = self._num_tail
        st = float(segment.get("start_time", 0.0))
        en = float(segment.get("end_time", st))
        tail["start_time"].append(st)
        tail["end_time"].append(en)
        tail["duration"].append(float(segment.get("duration", en - st)))
        tail["asr_success"].append(1 if segment.get("asr_success") else 0)
        tail["audio_success"].append(1 if segment.get("audio_success") else 0)
        tail["sample_rate"].append(int(segment.get("sample_rate", 0) or 0))
        self._text["segment_id"].tail.append(str(segment.get("segment_id", f"seg_{self._n}")))
        self._text["asr_text"].tail.append(str(segment.get("asr_text", "") or ""))
        self._text["audio_description"].tail.append(str(segment.get("audio_description", "") or ""))
        self._text["file"].tail.append(str(segment.get("file", file) or ""))
        self._text["extra_json"].tail.append(json.dumps(unknown, default=str) if unknown else "")
        audio = segment.get("audio_data")
        self._audio_tail.append(None if audio is None else np.asarray(audio))
        self._pending.append((segment.get("asr_embedding"), segment.get("audio_embedding")))
        self._n += 1

    def extend(self, segments: Iterable[Dict], file: Optional[str] = None) -> None:   # :797
        for s in segments:
            self.append(s, file=file)

    def clear(self) -> None:
        generation = self.generation
        self.__init__()
        self.generation = generation + 1

    # ---- embeddings: host -> device hand-over ---------------------------------------------------
    def bind_index(self, index) -> None:
        """The SegmentIndex that holds this table's embeddings (for lazy embedding fetches)."""
        self._index = index

    @property
    def n_pending(self) -> int:
        return len(self._pending)

    def pending_arrays(self):
        """Embeddings + flags of the rows appended since the last commit, as the arrays
        `SegmentIndex.append` takes (None embedding -> zero row, :640-641).  Nothing is forgotten
        yet: call `commit_pending(m)` once the device index has accepted them, so a failed append
        (NaN row -> ValueError, out of memory) leaves table and index in step and can be retried.
        Returns (first_row, asr, audio, flags)."""
        m = len(self._pending)
        row0 = self._pending_row0
        asr = np.zeros((m, DIM), dtype=np.float32)
        audio = np.zeros((m, DIM), dtype=np.float32)
        for i, (a, b) in enumerate(self._pending):
            if a is not None:
                asr[i] = _embedding_row(a)
            if b is not None:
                audio[i] = _embedding_row(b)
        flags = (self.column("asr_success", row0, row0 + m).astype(np.uint8)
                 | (self.column("audio_success", row0, row0 + m).astype(np.uint8) << 1))
        return row0, asr, audio, flags

    def commit_pending(self, m: int) -> None:
        """Forget the host copies of the first `m` pending rows: they live in HBM from here on."""
        del self._pending[:m]
        self._pending_row0 += m

    def drain_pending(self):
        """`pending_arrays()` + `commit_pending()` in one step (callers that cannot fail)."""
        out = self.pending_arrays()
        self.commit_pending(out[1].shape[0])
        return out

    # ---- row access ---------------------------------------------------------------------------
    def _num_at(self, name: str, i: int):
        return self._num[name][i] if i < self._n_base else self._num_tail[name][i - self._n_base]

    def column(self, name: str, r0: int = 0, r1: Optional[int] = None) -> np.ndarray:
        """Numeric column values of rows [r0, r1)."""
        r1 = self._n if r1 is None else r1
        dt = dict(NUMERIC_COLUMNS)[name]
        base = np.asarray(self._num[name][min(r0, self._n_base):min(r1, self._n_base)], dtype=dt)
        t0, t1 = max(r0 - self._n_base, 0), max(r1 - self._n_base, 0)
        if t1 > t0:
            return np.concatenate([base, np.asarray(self._num_tail[name][t0:t1], dtype=dt)])
        return base

    def success_flags(self, i: int):
        return bool(self._num_at("asr_success", i)), bool(self._num_at("audio_success", i))

    def _fast_views(self):
        """memoryviews of the loaded (base) columns: a scalar read through a memoryview returns a
        Python int / float in ~40 ns, through numpy (memmap) indexing in ~150 ns + a conversion; a
        result record reads 6 numbers and 10 text offsets."""
        num = {name: memoryview(np.ascontiguousarray(self._num[name])) for name, _ in NUMERIC_COLUMNS}
        text = {}
        for name in TEXT_COLUMNS:
            col = self._text[name]
            text[name] = (memoryview(np.ascontiguousarray(col.offsets, dtype=np.int64)), col.blob)
        self._fast = (num, text)
        return self._fast

    def _base_record(self, i: int) -> SegmentRecord:
        num, text = self._fast or self._fast_views()

        def txt(name):
            offs, blob = text[name]
            lo, hi = offs[i], offs[i + 1]
            return bytes(blob[lo:hi]).decode("utf-8") if hi > lo else ""
        fields = {
            "segment_id": txt("segment_id") or f"seg_{i}",
            "start_time": num["start_time"][i],
            "end_time": num["end_time"][i],
            "duration": num["duration"][i],
            "asr_text": txt("asr_text"),
            "asr_success": num["asr_success"][i] != 0,
            "audio_description": txt("audio_description"),
            "audio_success": num["audio_success"][i] != 0,
            "sample_rate": num["sample_rate"][i],
        }
        file = txt("file")
        if file:
            fields["file"] = file
        extra = txt("extra_json")
        if extra:
            fields.update(json.loads(extra))
        return SegmentRecord(self, i, fields)

    def record(self, i: int) -> SegmentRecord:
        if i < self._n_base:
            return self._base_record(i)
        fields = {
            "segment_id": self._text["segment_id"][i] or f"seg_{i}",
            "start_time": float(self._num_at("start_time", i)),
            "end_time": float(self._num_at("end_time", i)),
            "duration": float(self._num_at("duration", i)),
            "asr_text": self._text["asr_text"][i],
            "asr_success": bool(self._num_at("asr_success", i)),
            "audio_description": self._text["audio_description"][i],
            "audio_success": bool(self._num_at("audio_success", i)),
            "sample_rate": int(self._num_at("sample_rate", i)),
        }
        file = self._text["file"][i]
        if file:                                     # the reference's records have no such key
            fields["file"] = file
        extra = self._text["extra_json"][i]
        if extra:
            fields.update(json.loads(extra))
        return SegmentRecord(self, i, fields)

    def fetch(self, i: int, key: str):
        """Resolve a lazy field of row i."""
        if key == "audio_data":
            if i < self._n_base:
                lo, hi = int(self._audio_offsets[i]), int(self._audio_offsets[i + 1])
                if hi == lo or self._audio_blob is None:
                    return None
                return np.array(self._audio_blob[lo:hi])
            return self._audio_tail[i - self._n_base]
        corpus = 0 if key == "asr_embedding" else 1
        if i >= self._pending_row0:                                  # still on the host
            return self._pending[i - self._pending_row0][corpus]
        if self._index is None:
            raise KeyError(f"{key}: the table is not bound to a device index")
        # rows come back L2-normalised (the index stores normalize(Y), :646/:651 do it per call);
        # a zero row is a missing embedding
        local = i - int(self._index.row_base)
        row = self._index.read_rows(corpus, local, local + 1)[0]
        return row if np.any(row) else None

    # ---- persistence ----------------------------------------------------------------------------
    def save(self, path: str, audio: bool = True) -> None:
        """Write the columns to `path` (+ the audio samples to `path + '.audio'`)."""
        blocks = []                                  # (name, array)
        for name, dt in NUMERIC_COLUMNS:
            blocks.append((name, np.ascontiguousarray(self.column(name), dtype=dt)))
        for name in TEXT_COLUMNS:
            offs, blob = self._text[name].consolidated()
            blocks.append((name + ".offsets", np.ascontiguousarray(offs, dtype="<i8")))
            blocks.append((name + ".blob", np.ascontiguousarray(blob, dtype="u1")))
        tail_lens = np.fromiter((0 if a is None else a.size for a in self._audio_tail),
                                dtype=np.int64, count=len(self._audio_tail))
        base_offs = np.asarray(self._audio_offsets, dtype=np.int64)
        if not audio:
            audio_offs = np.zeros(self._n + 1, dtype=np.int64)
        else:
            audio_offs = np.concatenate([base_offs, int(base_offs[-1]) + np.cumsum(tail_lens)])
        blocks.append(("audio_offsets", np.ascontiguousarray(audio_offs, dtype="<i8")))

        directory = []
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(b"\0" * _HEADER.size)
            for name, arr in blocks:
                pos = (f.tell() + 63) // 64 * 64
                f.seek(pos)
                f.write(arr.tobytes())
                directory.append({"name": name, "dtype": arr.dtype.str, "offset": pos, "count": int(arr.size)})
            dir_off = (f.tell() + 63) // 64 * 64
            f.seek(dir_off)
            payload = json.dumps({"columns": directory}).encode("utf-8")
            f.write(payload)
            f.seek(0)
            f.write(_HEADER.pack(MAGIC, VERSION, 0, self._n, dir_off, len(payload)))
        if audio and int(audio_offs[-1]) > 0:
            atmp = path + ".audio.tmp"
            with open(atmp, "wb") as f:
                if self._audio_blob is not None:
                    step = 16 << 20                             # samples per write: the loaded blob may not fit in RAM
                    for lo in range(0, int(base_offs[-1]), step):
                        f.write(np.asarray(self._audio_blob[lo:min(lo + step, int(base_offs[-1]))], dtype="<f4").tobytes())
                for a in self._audio_tail:
                    if a is not None:
                        f.write(np.ascontiguousarray(a, dtype="<f4").tobytes())
            os.replace(atmp, path + ".audio")
        os.replace(tmp, path)

    @staticmethod
    def file_info(path: str) -> Dict:
        with open(path, "rb") as f:
            head = f.read(_HEADER.size)
            if len(head) != _HEADER.size:
                raise ValueError(f"{path}: too short for a segment table")
            magic, version, _, n, dir_off, dir_len = _HEADER.unpack(head)
            if magic != MAGIC:
                raise ValueError(f"{path}: not a segment table (bad magic)")
            if version != VERSION:
                raise ValueError(f"{path}: unsupported segment table version {version}")
            size = os.fstat(f.fileno()).st_size
            if dir_off + dir_len > size:
                raise ValueError(f"{path}: truncated (directory beyond end of file)")
            f.seek(dir_off)
            directory = json.loads(f.read(dir_len).decode("utf-8"))["columns"]
        for c in directory:
            if c["offset"] + c["count"] * np.dtype(c["dtype"]).itemsize > size:
                raise ValueError(f"{path}: truncated (column {c['name']} beyond end of file)")
        return {"n_rows": int(n), "columns": directory}

    @classmethod
    def load(cls, path: str) -> "SegmentTable":
        """Memory-map a saved table: no per-row work, pages come in as results are read."""
        info = cls.file_info(path)
        n = info["n_rows"]
        cols = {}
        for c in info["columns"]:
            cols[c["name"]] = (np.memmap(path, dtype=c["dtype"], mode="r", offset=c["offset"], shape=(c["count"],))
                               if c["count"] else np.zeros(0, dtype=c["dtype"]))
        t = cls()
        for name, _ in NUMERIC_COLUMNS:
            if len(cols[name]) != n:
                raise ValueError(f"{path}: column {name} has {len(cols[name])} rows, header says {n}")
            t._num[name] = cols[name]
        for name in TEXT_COLUMNS:
            offs = cols[name + ".offsets"]
            if len(offs) != n + 1:
                raise ValueError(f"{path}: column {name} has {len(offs) - 1} rows, header says {n}")
            t._text[name] = _TextColumn(offs, cols[name + ".blob"])
        t._audio_offsets = cols["audio_offsets"]
        if len(t._audio_offsets) != n + 1:
            raise ValueError(f"{path}: audio offsets do not match {n} rows")
        n_samples = int(t._audio_offsets[-1])
        if n_samples:
            apath = path + ".audio"
            if not os.path.exists(apath) or os.path.getsize(apath) < n_samples * 4:
                raise ValueError(f"{apath}: missing or truncated ({n_samples} samples expected)")
            t._audio_blob = np.memmap(apath, dtype="<f4", mode="r", shape=(n_samples,))
        t._n_base = t._n = n
        t._pending_row0 = n
        return t


def _embedding_row(e) -> np.ndarray:
    a = np.asarray(e, dtype=np.float32).reshape(-1)
    if a.shape[0] != DIM:
        raise ValueError(f"Incompatible dimension for X and Y matrices: X.shape[1] == {DIM} "
                         f"while Y.shape[1] == {a.shape[0]}")
    return a
