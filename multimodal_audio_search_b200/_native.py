"""ctypes binding of the C-ABI in include/cab.h (libcab.so, built in-tree by build.py).

There is deliberately no fallback: if the library is missing or no B200 is visible the product
path raises (`NativeLibraryMissing` / `CabError`) instead of computing anything on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcab.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "cab.h")

CAB_OK, CAB_ERR_INVALID, CAB_ERR_CUDA, CAB_ERR_NONFINITE, CAB_ERR_NO_DEVICE, CAB_ERR_NOMEM = range(6)
CAB_F32, CAB_BF16 = 0, 1
CAB_HOST, CAB_DEVICE = 0, 1
CAB_PATH_AUTO, CAB_PATH_GEMV, CAB_PATH_GEMM = 0, 1, 2
CAB_DIM, CAB_MAX_K, CAB_MAX_QUERIES = 384, 128, 4096
CANDIDATE_BYTES = 24
CAB_IPC_HANDLE_BYTES, CAB_MAX_WORLD = 64, 8


class NativeLibraryMissing(RuntimeError):
    pass


class CabError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"cab status {status}: {message}")
        self.status = status
        self.message = message


_p = C.c_void_p
_i64, _i32, _u32, _dbl = C.c_int64, C.c_int, C.c_uint32, C.c_double

# name -> (restype, argtypes); mirrors include/cab.h one to one (checked by tests/test_cabi_symbols.py)
SIGNATURES = {
    "cab_version": (_i32, []),
    "cab_status_string": (C.c_char_p, [_i32]),
    "cab_last_error": (C.c_char_p, [_p]),
    "cab_device_count": (_i32, []),
    "cab_index_create": (_i32, [_i32, _i32, _i64, _i32, C.POINTER(_p)]),
    "cab_index_destroy": (_i32, [_p]),
    "cab_index_reserve": (_i32, [_p, _i64]),
    "cab_index_size": (_i64, [_p]),
    "cab_index_capacity": (_i64, [_p]),
    "cab_index_dtype": (_i32, [_p]),
    "cab_index_device": (_i32, [_p]),
    "cab_index_set_row_base": (_i32, [_p, _i64]),
    "cab_index_row_base": (_i64, [_p]),
    "cab_index_clear": (_i32, [_p]),
    "cab_index_append": (_i32, [_p, _p, _p, _p, _i64, _i32, _p]),
    "cab_index_append_synth": (_i32, [_p, _u32, _i64, _i64, _i64, _i32, _i32, _i32, _p]),
    "cab_synth_queries": (_i32, [_i32, _u32, _i32, _i32, _p, _i32]),
    "cab_index_read_rows": (_i32, [_p, _i32, _i64, _i64, _p, _i32]),
    "cab_index_save": (_i32, [_p, C.c_char_p]),
    "cab_index_load": (_i32, [C.c_char_p, _i32, _i64, _i64, C.POINTER(_p)]),
    "cab_index_file_info": (_i32, [C.c_char_p, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64), C.POINTER(_i64)]),
    "cab_search": (_i32, [_p, _p, _i32, _p, _p, _i32, _i32, _dbl, _i32, _p, _p, _p, _p, _p, _p, _i32, _p]),
    "cab_search_candidates": (_i32, [_p, _p, _i32, _p, _p, _i32, _i32, _dbl, _i32, _p, _p]),
    "cab_merge_candidates": (_i32, [_p, _p, _i32, _i32, _i32, _p, _p, _dbl, _p, _p, _p, _p, _p, _p, _i32, _p]),
    "cab_peer_init": (_i32, [_p, _i32, _i32, _i32, _i32, _p]),
    "cab_peer_attach": (_i32, [_p, _p]),
    "cab_search_sharded": (_i32, [_p, _p, _i32, _p, _p, _i32, _i32, _dbl, _i32, _p, _p, _p, _p, _p, _p, _i32, _p]),
    "cab_index_exchange_stamps": (_i32, [_p, _p, _i32]),
    "cab_peer_snapshot": (_i32, [_p, _p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "cab_index_set_option": (_i32, [_p, C.c_char_p, _i64]),
    "cab_index_get_option": (_i64, [_p, C.c_char_p]),
    "cab_index_launch_count": (_i64, [_p]),
    "cab_index_last_scan_ms": (_dbl, [_p]),
    "cab_score_all": (_i32, [_p, _p, _i32, _i32, _p, _p, _i32, _p]),
    "cab_index_write_flags": (_i32, [_p, _i64, _i64, _p, _i32]),
    "cab_index_read_flags": (_i32, [_p, _i64, _i64, _p, _i32]),
}

_lib = None


def header_symbols() -> list[str]:
    """Every function declared in include/cab.h."""
    with open(HEADER_PATH) as f:
        return sorted(set(re.findall(r"^CAB_API[^(]*?\b(cab_[a-z0-9_]+)\s*\(", f.read(), flags=re.M)))


def lib() -> C.CDLL:
    """Load libcab.so (once).  Raises NativeLibraryMissing -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -m multimodal_audio_search_b200.build` "
            "(nvcc, sm_100a). The engine has no CPU fallback.")
    try:
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    except OSError as e:  # pragma: no cover
        raise NativeLibraryMissing(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return handle


TORCH_LIB_PATH = os.path.join(HERE, "libcab_torch.so")
_torch_ops = None


def torch_ops():
    """`torch.ops.cab` from the PyTorch extension (libcab_torch.so), loaded once."""
    global _torch_ops
    if _torch_ops is None:
        lib()                                   # libcab.so first (the extension links it)
        if not os.path.exists(TORCH_LIB_PATH):
            raise NativeLibraryMissing(
                f"{TORCH_LIB_PATH} not found: build it with `python -m multimodal_audio_search_b200.build`")
        import torch
        torch.ops.load_library(TORCH_LIB_PATH)
        _torch_ops = torch.ops.cab
    return _torch_ops


def check(status: int, handle=None):
    if status == CAB_OK:
        return
    msg = lib().cab_last_error(handle)
    text = msg.decode("utf-8", "replace") if msg else ""
    if not text:
        text = lib().cab_status_string(status).decode()
    if status == CAB_ERR_NONFINITE:
        raise ValueError(text)                 # what sklearn raises for the reference
    raise CabError(status, text)
