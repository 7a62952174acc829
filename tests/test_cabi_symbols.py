"""CPU-only: the C-ABI library builds, loads, and exports exactly what include/cab.h declares;
without a GPU every entry point fails loudly (no CPU fallback)."""
import ctypes as C

import numpy as np
import pytest

from multimodal_audio_search_b200 import _native as N


@pytest.fixture(scope="module")
def lib():
    from multimodal_audio_search_b200 import build
    build.build()
    return N.lib()


def test_header_and_binding_agree(lib):
    declared = N.header_symbols()
    assert len(declared) >= 20
    assert set(declared) == set(N.SIGNATURES), set(declared) ^ set(N.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cab.h but not exported"


def test_version_and_status_strings(lib):
    assert lib.cab_version() == 100
    assert b"no usable CUDA device" in lib.cab_status_string(N.CAB_ERR_NO_DEVICE)
    assert lib.cab_status_string(0) == b"ok"


def test_candidate_record_is_24_bytes():
    class Cand(C.Structure):
        _fields_ = [("index", C.c_int64), ("asr", C.c_float), ("audio", C.c_float),
                    ("flags", C.c_uint32), ("pad", C.c_uint32)]
    assert C.sizeof(Cand) == N.CANDIDATE_BYTES


def test_bad_arguments_rejected_before_touching_cuda(lib):
    h = C.c_void_p()
    assert lib.cab_index_create(128, N.CAB_F32, 0, 0, C.byref(h)) == N.CAB_ERR_INVALID
    assert b"384" in lib.cab_last_error(None)
    assert lib.cab_index_create(384, 7, 0, 0, C.byref(h)) == N.CAB_ERR_INVALID
    assert lib.cab_index_size(None) == -1


def test_no_cpu_fallback(lib):
    """On a box without a GPU the product path must raise, not compute."""
    if lib.cab_device_count() > 0:
        pytest.skip("a GPU is visible")
    from multimodal_audio_search_b200 import DualPipelineAudioSearch, SegmentIndex
    with pytest.raises(N.CabError) as e:
        SegmentIndex("fp32")
    assert e.value.status == N.CAB_ERR_NO_DEVICE

    class Emb:
        def encode(self, text):
            return np.ones(384, np.float32)
    eng = DualPipelineAudioSearch(text_embedder=Emb())
    assert eng.search_with_fusion("anything") == ([], {})          # empty library: reference behaviour
    eng.audio_segments.append({"asr_embedding": np.ones(384, np.float32), "audio_embedding": None,
                               "asr_success": True, "audio_success": False})
    with pytest.raises(N.CabError):
        eng.search_with_fusion("anything")


def test_torch_extension_registers_ops(lib):
    """The PyTorch extension (libcab_torch.so) builds against the installed torch and registers
    torch.ops.cab.* (no compute without a GPU)."""
    from multimodal_audio_search_b200 import build
    build.build_torch_extension()
    ops = N.torch_ops()
    for name in ("append", "search", "search_candidates", "merge_candidates"):
        assert hasattr(ops, name)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(Exception):      # CPU tensors are rejected, nothing is computed on the host
            ops.search(0, torch.zeros(1, 384), torch.ones(1), torch.ones(1), 10, 0.1, 0)
