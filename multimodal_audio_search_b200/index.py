"""`SegmentIndex`: the device-resident dual-corpus segment store + fused search, over the C-ABI.

Host numpy arrays go through the library's pinned staging (copies inside the call); CUDA torch
tensors are passed by pointer on torch's current stream and results come back as CUDA tensors.
PyTorch is used only as the owner of device memory / streams -- all arithmetic is in libcab.so.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _native as N

_DTYPES = {"fp32": N.CAB_F32, "f32": N.CAB_F32, "float32": N.CAB_F32,
           "bf16": N.CAB_BF16, "bfloat16": N.CAB_BF16}
_PATHS = {"auto": N.CAB_PATH_AUTO, "gemv": N.CAB_PATH_GEMV, "gemm": N.CAB_PATH_GEMM}


@dataclass
class SearchResult:
    """Per query: `count[q]` results, best first; slots beyond count hold index -1."""
    indices: "np.ndarray"      # int64  [Q, k] global segment index
    fusion: "np.ndarray"       # float64 [Q, k]
    asr_sim: "np.ndarray"      # float32 [Q, k]
    audio_sim: "np.ndarray"    # float32 [Q, k]
    flags: "np.ndarray"        # uint8   [Q, k]
    count: "np.ndarray"        # int32   [Q]


def _is_torch_cuda(x) -> bool:
    return hasattr(x, "is_cuda") and bool(x.is_cuda)


def _np_f32(x, cols=None):
    a = np.ascontiguousarray(x, dtype=np.float32)
    if cols is not None and (a.ndim != 2 or a.shape[1] != cols):
        raise ValueError(f"expected a [n x {cols}] float32 matrix, got shape {a.shape}")
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SegmentIndex:
    def __init__(self, dtype: str = "fp32", capacity: int = 0, device: int = 0, dim: int = N.CAB_DIM):
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        self._lib = N.lib()
        self._h = C.c_void_p()
        N.check(self._lib.cab_index_create(dim, _DTYPES[dtype], int(capacity), int(device), C.byref(self._h)))
        self.dtype = "bf16" if _DTYPES[dtype] == N.CAB_BF16 else "fp32"
        self.device = int(device)
        self.dim = dim

    # -- persistence ---------------------------------------------------------------------------
    def save(self, path: str):
        """Write the index (normalised rows + flags exactly as in HBM) to `path`."""
        N.check(self._lib.cab_index_save(self._h, os.fsencode(path)), self._h)

    @classmethod
    def load(cls, path: str, device: int = 0, rows: "tuple[int, int] | None" = None) -> "SegmentIndex":
        """Load a saved index (or the row range `rows=(r0, r1)` of it: a shard) onto `device`."""
        self = cls.__new__(cls)
        self._lib = N.lib()
        self._h = C.c_void_p()
        r0, r1 = (0, -1) if rows is None else rows
        N.check(self._lib.cab_index_load(os.fsencode(path), int(device), int(r0), int(r1), C.byref(self._h)))
        self.dtype = "bf16" if self._lib.cab_index_dtype(self._h) == N.CAB_BF16 else "fp32"
        self.device, self.dim = int(device), N.CAB_DIM
        return self

    @staticmethod
    def file_info(path: str) -> dict:
        """Header of a saved index without touching the GPU."""
        dim, dtype, n, base = C.c_int(), C.c_int(), C.c_int64(), C.c_int64()
        N.check(N.lib().cab_index_file_info(os.fsencode(path), C.byref(dim), C.byref(dtype), C.byref(n), C.byref(base)))
        return {"dim": dim.value, "dtype": "bf16" if dtype.value == N.CAB_BF16 else "fp32",
                "n_rows": n.value, "row_base": base.value}

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.cab_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self._lib.cab_index_size(self._h))

    @property
    def capacity(self) -> int:
        return int(self._lib.cab_index_capacity(self._h))

    @property
    def row_base(self) -> int:
        return int(self._lib.cab_index_row_base(self._h))

    @row_base.setter
    def row_base(self, v: int):
        N.check(self._lib.cab_index_set_row_base(self._h, int(v)), self._h)

    def reserve(self, rows: int):
        N.check(self._lib.cab_index_reserve(self._h, int(rows)), self._h)

    def clear(self):
        N.check(self._lib.cab_index_clear(self._h), self._h)

    def set_option(self, key: str, value: int):
        N.check(self._lib.cab_index_set_option(self._h, key.encode(), int(value)), self._h)

    def get_option(self, key: str) -> int:
        return int(self._lib.cab_index_get_option(self._h, key.encode()))

    @property
    def launch_count(self) -> int:
        return int(self._lib.cab_index_launch_count(self._h))

    def last_scan_ms(self) -> float:
        return float(self._lib.cab_index_last_scan_ms(self._h))

    def peer_snapshot(self) -> np.ndarray:
        """Diagnostic: this rank's exchange buffer as raw bytes (layout in include/cab.h)."""
        need = C.c_size_t()
        N.check(self._lib.cab_peer_snapshot(self._h, None, 0, C.byref(need)), self._h)
        out = np.zeros(need.value, dtype=np.uint8)
        N.check(self._lib.cab_peer_snapshot(self._h, _ptr(out), out.size, None), self._h)
        return out

    def enable_tensor_core_batches(self, on: bool = True):
        """fp32 index only: keep bf16 shadow copies of both corpora (+50 % HBM) so that batches of
        >= 4 queries (libraries of >= 64 K segments) are PRE-selected on the tensor cores, re-scored exactly from the fp32 rows and
        certified per query; a query whose top-k is not provably exact is re-run on the exact scan
        (`get_option("last_uncertified")` says how many of the last batch).  Results are those of
        the fp32 GEMV path; throughput is the bf16 tensor-core path's."""
        self.set_option("tensor_core_shadow", int(bool(on)))

    def exchange_stamps(self, max_rows: int = 64) -> np.ndarray:
        """Option "stamp_exchange": uint64 [n, 6] %globaltimer ns of the last sharded searches --
        {scan complete, best k selected, winners re-scored, own flag raised, all ranks' flags seen,
        results written} (oldest first)."""
        out = np.zeros((max_rows, 8), dtype=np.uint64)
        n = int(self._lib.cab_index_exchange_stamps(self._h, _ptr(out), int(max_rows)))
        return out[:n, :6]

    def _stream(self):
        """torch's current stream for the C-ABI.  NULL means "the handle's own stream" there, and
        torch's default stream IS the NULL handle, so it is passed as cudaStreamLegacy (0x1)."""
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream or 1)

    # -- ingest --------------------------------------------------------------------------------
    def append(self, asr_rows, audio_rows, flags=None):
        """Append segments.  Rows: [n x 384] float32 (numpy, or CUDA torch tensors), raw; `None`
        for a whole corpus means "no embedding" for every row.  flags: uint8 [n] (bit0 asr_success,
        bit1 audio_success) or None (= 3)."""
        if _is_torch_cuda(asr_rows) or _is_torch_cuda(audio_rows):
            import torch
            ts = [t for t in (asr_rows, audio_rows) if t is not None]
            n = ts[0].shape[0]
            for t in ts:
                if t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] != self.dim or not t.is_contiguous():
                    raise ValueError("device rows must be contiguous float32 [n x 384]")
            f = None
            if flags is not None:
                f = flags if _is_torch_cuda(flags) else torch.as_tensor(np.asarray(flags, np.uint8)).cuda(self.device)
                if f.dtype != torch.uint8 or f.numel() != n:
                    raise ValueError("flags must be uint8 [n]")
            N.torch_ops().append(self._h.value, asr_rows, audio_rows, f)
            return
        a = None if asr_rows is None else _np_f32(asr_rows, self.dim)
        b = None if audio_rows is None else _np_f32(audio_rows, self.dim)
        if a is None and b is None:
            raise ValueError("at least one corpus must be given")
        n = (a if a is not None else b).shape[0]
        if a is not None and b is not None and a.shape[0] != b.shape[0]:
            raise ValueError("asr_rows and audio_rows differ in length")
        f = None
        if flags is not None:
            f = np.ascontiguousarray(flags, dtype=np.uint8)
            if f.shape != (n,):
                raise ValueError("flags must be uint8 [n]")
        N.check(self._lib.cab_index_append(self._h, _ptr(a), _ptr(b), _ptr(f), n, N.CAB_HOST, None), self._h)

    def append_synth(self, seed: int, n_total: int, r0: int = 0, r1: int | None = None,
                     n_queries: int = 1, plants: int = 0, partial: bool = False, mode: str = "planted"):
        """Generate global rows [r0, r1) of the synthetic library on the device (synth.py twin);
        `mode`: "planted" | "ascending" | "clustered" (synth.MODES)."""
        from .synth import MODES
        r1 = n_total if r1 is None else r1
        N.check(self._lib.cab_index_append_synth(self._h, seed & 0xFFFFFFFF, n_total, r0, r1,
                                                 n_queries, plants, int(bool(partial)) | (MODES[mode] << 8), None), self._h)

    def read_rows(self, corpus: int, r0: int, r1: int) -> np.ndarray:
        out = np.empty((r1 - r0, self.dim), dtype=np.float32)
        N.check(self._lib.cab_index_read_rows(self._h, corpus, r0, r1, _ptr(out), N.CAB_HOST), self._h)
        return out

    def read_flags(self, r0: int = 0, r1: "int | None" = None) -> np.ndarray:
        """Flag bytes of rows [r0, r1): bit0 asr_success, bit1 audio_success, bits 2-3 weight class."""
        r1 = len(self) if r1 is None else r1
        out = np.empty(max(r1 - r0, 0), dtype=np.uint8)
        N.check(self._lib.cab_index_read_flags(self._h, int(r0), int(r1), _ptr(out), N.CAB_HOST), self._h)
        return out

    def write_flags(self, flags, r0: int = 0):
        """Overwrite the flag bytes of rows [r0, r0 + len(flags)) (numpy uint8 or CUDA uint8 tensor)."""
        if _is_torch_cuda(flags):
            import torch
            if flags.dtype != torch.uint8 or flags.dim() != 1 or not flags.is_contiguous():
                raise ValueError("flags must be a contiguous uint8 vector")
            torch.cuda.current_stream(self.device).synchronize()      # the copy runs on the handle's own stream
            N.check(self._lib.cab_index_write_flags(self._h, int(r0), int(r0) + flags.numel(),
                                                    C.c_void_p(flags.data_ptr()), N.CAB_DEVICE), self._h)
            return
        f = np.ascontiguousarray(flags, dtype=np.uint8).reshape(-1)
        N.check(self._lib.cab_index_write_flags(self._h, int(r0), int(r0) + f.size, _ptr(f), N.CAB_HOST), self._h)

    def set_weight_classes(self, classes, r0: int = 0):
        """Assign the 2-bit weight class (0..3) of rows [r0, r0 + len(classes)) for `score_all`,
        keeping their success bits."""
        c = np.ascontiguousarray(classes).reshape(-1)
        if c.size and (c.min() < 0 or c.max() > 3):
            raise ValueError("weight classes must be in 0..3")
        cur = self.read_flags(r0, r0 + c.size)
        self.write_flags((cur & 3) | (c.astype(np.uint8) << 2), r0)

    # -- search --------------------------------------------------------------------------------
    @staticmethod
    def _weights(w_asr, w_audio, nq):
        """Owned float64 [nq] arrays (scalars broadcast)."""
        if nq == 1 and isinstance(w_asr, float) and isinstance(w_audio, float):      # the drop-in's call: two Python floats
            return np.array([w_asr]), np.array([w_audio])
        wa = np.asarray(w_asr, dtype=np.float64).reshape(-1)
        wb = np.asarray(w_audio, dtype=np.float64).reshape(-1)
        if wa.size != nq:
            wa = np.full(nq, wa[0]) if wa.size == 1 else wa
        if wb.size != nq:
            wb = np.full(nq, wb[0]) if wb.size == 1 else wb
        if wa.size != nq or wb.size != nq:
            raise ValueError("one weight per query (or a scalar) expected")
        return np.ascontiguousarray(wa), np.ascontiguousarray(wb)

    def search(self, queries, w_asr=0.5, w_audio=0.5, k: int = 10, threshold: float = 0.1,
               path: str = "auto") -> SearchResult:
        """Fused dual-corpus top-k.  numpy queries -> numpy results (host staging + copies inside
        the call); CUDA torch queries -> CUDA torch results on torch's current stream."""
        if _is_torch_cuda(queries):
            return self._search_device(queries, w_asr, w_audio, k, threshold, path)
        q = queries
        if not (isinstance(q, np.ndarray) and q.dtype == np.float32 and q.flags.c_contiguous):
            q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"expected [n x {self.dim}] float32 queries, got shape {q.shape}")
        nq = q.shape[0]
        wa, wb = self._weights(w_asr, w_audio, nq)
        oi, of = np.empty((nq, k), np.int64), np.empty((nq, k), np.float64)
        oa, ob = np.empty((nq, k), np.float32), np.empty((nq, k), np.float32)
        ofl, oc = np.empty((nq, k), np.uint8), np.empty((nq,), np.int32)
        rc = self._lib.cab_search(self._h, q.ctypes.data, N.CAB_HOST, wa.ctypes.data, wb.ctypes.data, nq, k,
                                  threshold, _PATHS[path], oi.ctypes.data, of.ctypes.data, oa.ctypes.data,
                                  ob.ctypes.data, ofl.ctypes.data, oc.ctypes.data, N.CAB_HOST, None)
        if rc:
            N.check(rc, self._h)
        return SearchResult(oi, of, oa, ob, ofl, oc)

    def _search_device(self, queries, w_asr, w_audio, k, threshold, path) -> SearchResult:
        """CUDA tensors in, CUDA tensors out, through the PyTorch extension (torch.ops.cab.search)."""
        import torch
        q = queries if queries.dim() == 2 else queries.unsqueeze(0)
        wa, wb = self._weights(w_asr, w_audio, q.shape[0])
        out = N.torch_ops().search(self._h.value, q, torch.from_numpy(wa), torch.from_numpy(wb), int(k),
                                   float(threshold), _PATHS[path])
        return SearchResult(*out)

    # -- legacy modes: every segment's fused score ---------------------------------------------
    def score_all(self, queries, class_weights, out=None):
        """All-N fused similarities, no threshold / top-k (the earlier engine's
        `UnifiedAudioSearch.search`, previous_iterations/streamlit_app.py:173-223).
        class_weights: [4][2] (or [Q][4][2]) = {w_asr, w_audio} per row weight class (bits 2-3 of
        the row's flag byte).  numpy queries -> numpy float32 [Q, N]; CUDA tensor -> CUDA tensor.
        `out`: optional C-contiguous float32 [Q, N] numpy array to fill (host results only); a
        view of page-locked memory (`pinned_scores`) makes the copy back run at PCIe speed."""
        n = len(self)
        dev_in = _is_torch_cuda(queries)
        if dev_in:
            import torch
            q = queries if queries.dim() == 2 else queries.unsqueeze(0)
            if q.dtype != torch.float32 or q.shape[1] != self.dim or not q.is_contiguous():
                raise ValueError(f"expected contiguous float32 [n x {self.dim}] queries")
            nq = q.shape[0]
        else:
            q = _np_f32(np.atleast_2d(queries), self.dim)
            nq = q.shape[0]
        cw = np.asarray(class_weights, dtype=np.float32)
        if cw.shape == (4, 2):
            cw = np.broadcast_to(cw, (nq, 4, 2))
        if cw.shape != (nq, 4, 2):
            raise ValueError("class_weights must be [4][2] or [n_queries][4][2]")
        cw = np.ascontiguousarray(cw)
        if dev_in:
            out = torch.empty((nq, n), dtype=torch.float32, device=q.device)
            N.check(self._lib.cab_score_all(self._h, C.c_void_p(q.data_ptr()), N.CAB_DEVICE, nq, _ptr(cw),
                                            C.c_void_p(out.data_ptr()), N.CAB_DEVICE, self._stream()), self._h)
            return out
        if out is None:
            out = np.empty((nq, n), dtype=np.float32)
        elif not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == (nq, n)
                  and out.flags.c_contiguous and out.flags.writeable):
            raise ValueError(f"out must be a writable C-contiguous float32 array of shape {(nq, n)}")
        N.check(self._lib.cab_score_all(self._h, _ptr(q), N.CAB_HOST, nq, _ptr(cw), _ptr(out),
                                        N.CAB_HOST, None), self._h)
        return out

    def pinned_scores(self, n_queries: int = 1) -> np.ndarray:
        """A page-locked float32 [n_queries, len(self)] numpy view to pass as `score_all(out=)`;
        the buffer is owned by the index and reused (grown geometrically) between calls."""
        import torch
        need = int(n_queries) * len(self)
        buf = getattr(self, "_pinned", None)
        if buf is None or buf.numel() < need:
            buf = torch.empty(max(need, 2 * (buf.numel() if buf is not None else 0), 1), dtype=torch.float32,
                              pin_memory=True)
            self._pinned = buf
        return buf[:need].view(int(n_queries), len(self)).numpy()

    # -- sharded search (corpus split by segment over ranks) -----------------------------------
    def search_candidates(self, queries, w_asr=0.5, w_audio=0.5, k: int = 10,
                          threshold: float = 0.1, path: str = "auto"):
        """Local top-k of this shard as packed cab_candidate records: CUDA uint8 [Q, k, 24]."""
        import torch
        if _is_torch_cuda(queries):
            q = queries if queries.dim() == 2 else queries.unsqueeze(0)
            wa, wb = self._weights(w_asr, w_audio, q.shape[0])
            return N.torch_ops().search_candidates(self._h.value, q, torch.from_numpy(wa), torch.from_numpy(wb),
                                                   int(k), float(threshold), _PATHS[path])
        q = _np_f32(np.atleast_2d(queries), self.dim)
        qp, loc, nq = _ptr(q), N.CAB_HOST, q.shape[0]
        wa, wb = self._weights(w_asr, w_audio, nq)
        out = torch.empty((nq, k, N.CANDIDATE_BYTES), dtype=torch.uint8, device=f"cuda:{self.device}")
        N.check(self._lib.cab_search_candidates(self._h, qp, loc, _ptr(wa), _ptr(wb), nq, k,
                                                float(threshold), _PATHS[path],
                                                C.c_void_p(out.data_ptr()), self._stream()), self._h)
        return out

    def merge_candidates(self, gathered, w_asr=0.5, w_audio=0.5, k: int = 10,
                         threshold: float = 0.1, to_host: bool = True) -> SearchResult:
        """Merge candidate blocks of all shards: CUDA uint8 [world, Q, k, 24] -> final top-k.
        w_asr = w_audio = None reuses the weights staged by the preceding search_candidates."""
        import torch
        world, nq = gathered.shape[0], gathered.shape[1]
        if gathered.shape[2] != k or gathered.shape[3] != N.CANDIDATE_BYTES or not gathered.is_contiguous():
            raise ValueError("gathered must be contiguous uint8 [world, Q, k, 24]")
        wa, wb = (None, None) if w_asr is None else self._weights(w_asr, w_audio, nq)
        if to_host:
            out = SearchResult(np.empty((nq, k), np.int64), np.empty((nq, k), np.float64),
                               np.empty((nq, k), np.float32), np.empty((nq, k), np.float32),
                               np.empty((nq, k), np.uint8), np.empty((nq,), np.int32))
            p, loc = _ptr, N.CAB_HOST
        else:
            dev = gathered.device
            out = SearchResult(torch.empty((nq, k), dtype=torch.int64, device=dev),
                               torch.empty((nq, k), dtype=torch.float64, device=dev),
                               torch.empty((nq, k), dtype=torch.float32, device=dev),
                               torch.empty((nq, k), dtype=torch.float32, device=dev),
                               torch.empty((nq, k), dtype=torch.uint8, device=dev),
                               torch.empty((nq,), dtype=torch.int32, device=dev))
            p, loc = (lambda t: C.c_void_p(t.data_ptr())), N.CAB_DEVICE
        N.check(self._lib.cab_merge_candidates(self._h, C.c_void_p(gathered.data_ptr()), world, nq, k,
                                               _ptr(wa), _ptr(wb), float(threshold), p(out.indices),
                                               p(out.fusion), p(out.asr_sim), p(out.audio_sim),
                                               p(out.flags), p(out.count), loc, self._stream()), self._h)
        return out


    # -- sharded search with the exchange fused into the kernels (NVLink peer memory) -----------
    def peer_init(self, rank: int, world: int, max_queries: int = 256, max_k: int = N.CAB_MAX_K) -> bytes:
        """Allocate this rank's exchange buffer; returns its 64-byte CUDA IPC handle."""
        buf = C.create_string_buffer(N.CAB_IPC_HANDLE_BYTES)
        N.check(self._lib.cab_peer_init(self._h, int(rank), int(world), int(max_queries), int(max_k), buf), self._h)
        return buf.raw

    def peer_attach(self, all_handles: bytes):
        """Map every rank's exchange buffer (handles concatenated in rank order)."""
        N.check(self._lib.cab_peer_attach(self._h, C.c_char_p(all_handles)), self._h)

    def search_sharded(self, queries, w_asr=0.5, w_audio=0.5, k: int = 10, threshold: float = 0.1,
                       path: str = "auto", to_host: bool = True) -> SearchResult:
        """Collective: every rank calls it with the same shapes; the per-shard top-k travel over
        peer memory inside the finalize kernel; every rank gets the merged result."""
        import torch
        if _is_torch_cuda(queries):
            q = queries if queries.dim() == 2 else queries.unsqueeze(0)
            qp, qloc, nq, stream = C.c_void_p(q.data_ptr()), N.CAB_DEVICE, q.shape[0], self._stream()
        else:
            q = _np_f32(np.atleast_2d(queries), self.dim)
            qp, qloc, nq, stream = _ptr(q), N.CAB_HOST, q.shape[0], (None if to_host else self._stream())
        wa, wb = self._weights(w_asr, w_audio, nq)
        if to_host:
            out = SearchResult(np.empty((nq, k), np.int64), np.empty((nq, k), np.float64),
                               np.empty((nq, k), np.float32), np.empty((nq, k), np.float32),
                               np.empty((nq, k), np.uint8), np.empty((nq,), np.int32))
            p, loc = _ptr, N.CAB_HOST
        else:
            dev = f"cuda:{self.device}"
            out = SearchResult(torch.empty((nq, k), dtype=torch.int64, device=dev),
                               torch.empty((nq, k), dtype=torch.float64, device=dev),
                               torch.empty((nq, k), dtype=torch.float32, device=dev),
                               torch.empty((nq, k), dtype=torch.float32, device=dev),
                               torch.empty((nq, k), dtype=torch.uint8, device=dev),
                               torch.empty((nq,), dtype=torch.int32, device=dev))
            p, loc = (lambda t: C.c_void_p(t.data_ptr())), N.CAB_DEVICE
        N.check(self._lib.cab_search_sharded(self._h, qp, qloc, _ptr(wa), _ptr(wb), nq, k, float(threshold),
                                             _PATHS[path], p(out.indices), p(out.fusion), p(out.asr_sim),
                                             p(out.audio_sim), p(out.flags), p(out.count), loc, stream), self._h)
        return out


def synth_queries(seed: int, q0: int, q1: int, device: int = 0) -> np.ndarray:
    """Raw synthetic query vectors generated by the device twin of synth.raw_queries."""
    out = np.empty((q1 - q0, N.CAB_DIM), dtype=np.float32)
    N.check(N.lib().cab_synth_queries(device, seed & 0xFFFFFFFF, q0, q1, _ptr(out), N.CAB_HOST))
    return out
