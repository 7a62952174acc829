"""Bit-reproducible synthetic segment libraries (host / numpy side).

BASELINE.json's configs are all "synthetic unit vectors" libraries; the reference has no
fixtures.  Every element here is a small INTEGER stored as fp32, produced by
a counter-based 32-bit hash of (seed, stream, global_row, column), so

  * the numpy generator below and the CUDA generator (csrc/cab_ingest.cu, `cab_index_append_synth`) emit
    bit-identical rows on any machine -- no libm, no float rounding in the generator;
  * any row range [r0, r1) can be generated independently, so a corpus sharded over 1/2/4/8 GPUs
    is the same corpus;
  * golden fixtures only need (seed, shape) + the reference's outputs.

Rows are raw (not unit length): the engine L2-normalises on ingest exactly like the reference
re-normalises on every call (sklearn `normalize`).  Element
distribution: sum of four hash bytes minus 510 (Irwin-Hall n=4, ~Gaussian, sigma ~= 147.8),
so cosines between unrelated rows are ~N(0, 1/384) like isotropic unit vectors.

Planted neighbours: for query `qi`, plant `t`, the global plant number
g = qi*plants + t overwrites row p(g) = (base + g*stride) mod N (stride coprime to N, hence
collision-free) with  m*q_raw + n*noise_raw  for small integers m, n in [1, 8]  (cosine to the
query ~= m/sqrt(m^2+n^2), in [0.12, 0.99]); g%3 selects ASR-only / audio-only / both corpora.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

DIM = 384
STREAM_ASR, STREAM_AUDIO, STREAM_QUERY, STREAM_FLAGS = 0, 1, 2, 3
STREAM_CLUSTER, STREAM_PLANE, STREAM_CENTRE = 4, 5, 8       # 5 + corpus: second axis of the ascending plane

# Corpus distributions: "planted" = isotropic noise + planted neighbours
# (the headline data); two stress distributions for the top-k pruning:
#   "ascending": row r = alpha(r) * D + sqrt(1 - alpha(r)^2) * E_corpus with alpha rising linearly
#       from 0.15 to 0.95 over the GLOBAL row range (D = raw query 0, E = one fixed vector per
#       corpus): every row scores a hair above its predecessor for any query near D, so the
#       running k-th best never rejects anything -- the adversarial order for a pruning scan.
#       fp32, one correctly rounded operation at a time => numpy and CUDA agree bit for bit.
#   "clustered": row r = 3 * centre[c(r)] + 2 * noise_r with 1024 shared centres (anisotropic,
#       MiniLM-like): ~N/1024 rows score far above the threshold for a query near a centre.
MODES = {"planted": 0, "ascending": 1, "clustered": 2}
N_CLUSTERS = 1024
ASC_ALPHA0, ASC_SPAN = 0.15, 0.8

_U32 = np.uint32
_GOLD = 0x9E3779B9
_C1 = 0x85EBCA6B
_COLMUL = 0x9E3779B1
_PLANT_BASE_SALT = 0xABCD1234
_PLANT_MIX_SALT = 0x51ED270B

FLAG_ASR, FLAG_AUDIO = 1, 2


def _mix32(x: np.ndarray) -> np.ndarray:
    """lowbias32 integer finaliser; x is a uint32 array (wraps mod 2^32)."""
    x = x.astype(_U32, copy=True)
    x ^= x >> _U32(16)
    x *= _U32(0x7FEB352D)
    x ^= x >> _U32(15)
    x *= _U32(0x846CA68B)
    x ^= x >> _U32(16)
    return x


def _mix32_scalar(x: int) -> int:
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def stream_key(seed: int, stream: int) -> int:
    return _mix32_scalar((seed & 0xFFFFFFFF) ^ ((stream * _GOLD + _C1) & 0xFFFFFFFF))


def _row_keys(seed: int, stream: int, rows: np.ndarray) -> np.ndarray:
    return _mix32(_U32(stream_key(seed, stream)) ^ rows.astype(_U32))


def raw_rows(seed: int, stream: int, r0: int, r1: int, dim: int = DIM) -> np.ndarray:
    """Noise rows [r0, r1) of a stream as integer-valued fp32, shape (r1-r0, dim)."""
    rows = np.arange(r0, r1, dtype=np.uint64).astype(_U32)
    rk = _row_keys(seed, stream, rows)
    cols = (np.arange(1, dim + 1, dtype=np.uint64) * _COLMUL).astype(_U32)
    h = _mix32(rk[:, None] + cols[None, :])
    s = (h & _U32(0xFF)) + ((h >> _U32(8)) & _U32(0xFF)) + ((h >> _U32(16)) & _U32(0xFF)) + (h >> _U32(24))
    return s.astype(np.float32) - np.float32(510.0)


@dataclass(frozen=True)
class PlantSpec:
    """Collision-free placement of planted neighbours in a library of n_rows global rows."""
    n_rows: int
    n_queries: int
    plants: int          # planted rows per query
    base: int
    stride: int
    inv_stride: int      # stride^-1 mod n_rows

    @property
    def total(self) -> int:
        return self.n_queries * self.plants


def plant_spec(seed: int, n_rows: int, n_queries: int, plants: int) -> PlantSpec:
    if n_queries * plants > n_rows:
        raise ValueError("more planted rows than library rows")
    if n_rows <= 1 or n_queries * plants == 0:
        return PlantSpec(n_rows, n_queries, 0 if n_rows <= 1 else plants, 0, 1, 0 if n_rows <= 1 else 1)
    stride = int(n_rows * 0.6180339887) | 1
    while math.gcd(stride, n_rows) != 1:
        stride += 2
    stride %= n_rows
    if stride == 0:
        stride = 1
    base = _mix32_scalar((seed & 0xFFFFFFFF) ^ _PLANT_BASE_SALT) % n_rows
    return PlantSpec(n_rows, n_queries, plants, base, stride, pow(stride, -1, n_rows))


def plant_of_rows(spec: PlantSpec, rows: np.ndarray) -> np.ndarray:
    """Global plant number g for each global row, or -1 where the row is not planted."""
    if spec.total == 0:
        return np.full(rows.shape, -1, dtype=np.int64)
    n = np.uint64(spec.n_rows)
    d = (rows.astype(np.uint64) + n - np.uint64(spec.base)) % n
    g = (d * np.uint64(spec.inv_stride)) % n          # < n^2 < 2^64 for n <= 2^32
    g = g.astype(np.int64)
    return np.where(g < spec.total, g, -1)


def plant_rows_of_query(spec: PlantSpec, qi: int) -> np.ndarray:
    g = np.arange(qi * spec.plants, (qi + 1) * spec.plants, dtype=np.int64)
    return (spec.base + g * spec.stride) % spec.n_rows


def _plant_coeffs(seed: int, g: np.ndarray, stream: int):
    hm = _mix32(_U32((seed & 0xFFFFFFFF) ^ _PLANT_MIX_SALT) ^ (g.astype(np.uint64) * _COLMUL).astype(_U32))
    sh = _U32(8 * stream)
    m = 1 + ((hm >> sh) & _U32(7)).astype(np.int64)
    n = 1 + ((hm >> (sh + _U32(4))) & _U32(7)).astype(np.int64)
    return m, n


def raw_queries(seed: int, q0: int, q1: int, dim: int = DIM) -> np.ndarray:
    return raw_rows(seed, STREAM_QUERY, q0, q1, dim)


def _rows_at(seed: int, stream: int, rows: np.ndarray, dim: int = DIM) -> np.ndarray:
    """raw_rows for an arbitrary array of global row numbers."""
    rk = _row_keys(seed, stream, np.asarray(rows, dtype=np.uint64).astype(_U32))
    cols = (np.arange(1, dim + 1, dtype=np.uint64) * _COLMUL).astype(_U32)
    h = _mix32(rk[:, None] + cols[None, :])
    s = (h & _U32(0xFF)) + ((h >> _U32(8)) & _U32(0xFF)) + ((h >> _U32(16)) & _U32(0xFF)) + (h >> _U32(24))
    return s.astype(np.float32) - np.float32(510.0)


def cluster_of_rows(seed: int, rows: np.ndarray) -> np.ndarray:
    """Cluster id (0..1023) of each global row of a "clustered" library (same for both corpora)."""
    return (_row_keys(seed, STREAM_CLUSTER, np.asarray(rows, dtype=np.uint64).astype(_U32)) % _U32(N_CLUSTERS)).astype(np.int64)


def ascending_alpha(rows: np.ndarray, n_rows: int):
    """(alpha, beta) of the "ascending" library, float32, one rounded operation at a time."""
    step = np.float32(ASC_SPAN) / np.float32(max(n_rows - 1, 1))
    alpha = np.float32(ASC_ALPHA0) + np.asarray(rows, dtype=np.uint64).astype(np.float32) * step
    beta = np.sqrt(np.float32(1.0) - alpha * alpha)
    return alpha.astype(np.float32), beta.astype(np.float32)


def corpus_rows_at(seed: int, stream: int, rows: np.ndarray, spec: PlantSpec | None = None,
                   mode: str = "planted", n_rows: int | None = None, dim: int = DIM) -> np.ndarray:
    """The given global rows of the ASR (stream 0) or audio (stream 1) corpus."""
    rows = np.asarray(rows, dtype=np.int64)
    if mode == "ascending":
        alpha, beta = ascending_alpha(rows, spec.n_rows if n_rows is None else n_rows)
        d = raw_queries(seed, 0, 1, dim)[0]
        e = raw_rows(seed, STREAM_PLANE + stream, 0, 1, dim)[0]
        return alpha[:, None] * d[None, :] + beta[:, None] * e[None, :]
    x = _rows_at(seed, stream, rows, dim)
    if mode == "clustered":
        centres = _rows_at(seed, STREAM_CENTRE, cluster_of_rows(seed, rows), dim)
        return np.float32(3.0) * centres + np.float32(2.0) * x
    if mode != "planted":
        raise ValueError(f"unknown distribution {mode!r}")
    if spec is None or spec.total == 0:
        return x
    g = plant_of_rows(spec, rows)
    kind = g % 3                       # 0: ASR only, 1: audio only, 2: both
    hit = (g >= 0) & ((kind == 2) | (kind == stream))
    if hit.any():
        gi = g[hit]
        m, n = _plant_coeffs(seed, gi, stream)
        qi = gi // spec.plants
        uq, inv = np.unique(qi, return_inverse=True)
        qraw = np.concatenate([raw_queries(seed, int(u), int(u) + 1, dim) for u in uq], axis=0)
        x[hit] = m[:, None].astype(np.float32) * qraw[inv] + n[:, None].astype(np.float32) * x[hit]
    return x


def corpus_rows(seed: int, stream: int, r0: int, r1: int, spec: PlantSpec | None = None,
                dim: int = DIM, mode: str = "planted", n_rows: int | None = None) -> np.ndarray:
    """Rows [r0, r1) of the ASR (stream 0) or audio (stream 1) corpus, with plants applied."""
    return corpus_rows_at(seed, stream, np.arange(r0, r1, dtype=np.int64), spec, mode, n_rows, dim)


def bench_queries(seed: int, mode: str, q0: int, q1: int, dim: int = DIM) -> np.ndarray:
    """Raw query vectors that go with a distribution: the plain synthetic queries for "planted";
    4 * D + noise_i for "ascending" (cosine 0.97 to the ascent direction); 2 * centre[i % 1024] +
    noise_i for "clustered".  Integer valued, exact."""
    q = raw_queries(seed, q0, q1, dim)
    if mode == "ascending":
        return np.float32(4.0) * raw_queries(seed, 0, 1, dim) + q
    if mode == "clustered":
        return np.float32(2.0) * _rows_at(seed, STREAM_CENTRE, np.arange(q0, q1) % N_CLUSTERS, dim) + q
    return q


def row_flags_at(seed: int, rows: np.ndarray, partial: bool) -> np.ndarray:
    """Per-row pipeline-success flags (bit0 ASR, bit1 audio).  `partial`: ~10 % ASR-only and
    ~10 % audio-only rows (SURVEY.md section 8(d) validity-mask variant); otherwise all rows 3."""
    n = len(rows)
    if not partial:
        return np.full(n, FLAG_ASR | FLAG_AUDIO, dtype=np.uint8)
    hv = _mix32(_row_keys(seed, STREAM_FLAGS, np.asarray(rows, dtype=np.uint64).astype(_U32))) % _U32(10)
    f = np.full(n, FLAG_ASR | FLAG_AUDIO, dtype=np.uint8)
    f[hv == 0] = FLAG_ASR
    f[hv == 1] = FLAG_AUDIO
    return f


def row_flags(seed: int, r0: int, r1: int, partial: bool) -> np.ndarray:
    return row_flags_at(seed, np.arange(r0, r1, dtype=np.int64), partial)


def library(seed: int, n_rows: int, n_queries: int = 1, plants: int = 0, partial: bool = False,
            r0: int = 0, r1: int | None = None, dim: int = DIM, mode: str = "planted", rows=None):
    """(asr_rows, audio_rows, flags, spec) for global rows [r0, r1) -- or the explicit global row
    list `rows` -- of an n_rows library.
    A pipeline whose flag bit is clear has NO embedding in the reference (`None`,
    audio_search.py:344/350 -> similarity 0.0 at :640-641): its row is all zeros here."""
    r1 = n_rows if r1 is None else r1
    spec = plant_spec(seed, n_rows, n_queries, plants if mode == "planted" else 0)
    idx = np.arange(r0, r1, dtype=np.int64) if rows is None else np.asarray(rows, dtype=np.int64)
    a = corpus_rows_at(seed, STREAM_ASR, idx, spec, mode, n_rows, dim)
    b = corpus_rows_at(seed, STREAM_AUDIO, idx, spec, mode, n_rows, dim)
    f = row_flags_at(seed, idx, partial)
    a[(f & FLAG_ASR) == 0] = 0.0
    b[(f & FLAG_AUDIO) == 0] = 0.0
    return a, b, f, spec
