"""CPU-only: the C-ABI library builds, loads, and exports exactly what include/cab.h declares;
without a GPU every entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os

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


def test_header_is_plain_c_and_a_c_program_links(lib, tmp_path):
    """include/cab.h must be usable from C (the boundary has no C++ or torch types): compile it as
    C99 with -pedantic, and link a C program against libcab.so that calls only entry points which
    do not need a GPU."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = os.path.join(root, "include", "cab.h")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c", header], check=True)
    src = tmp_path / "main.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "cab.h"\n'
        "int main(void) {\n"
        "    cab_index *idx = NULL;\n"
        "    if (cab_version() != CAB_VERSION) return 1;\n"
        "    if (sizeof(cab_candidate) != 24) return 2;\n"
        "    if (cab_index_create(128, CAB_F32, 0, 0, &idx) != CAB_ERR_INVALID || idx != NULL) return 3;   /* wrong dim */\n"
        "    if (!strstr(cab_last_error(NULL), \"dim\")) return 4;\n"
        "    if (cab_search(NULL, NULL, 0, NULL, NULL, 1, 10, 0.1, 0, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL) != CAB_ERR_INVALID) return 5;\n"
        "    if (cab_score_all(NULL, NULL, 0, 1, NULL, NULL, 0, NULL) != CAB_ERR_INVALID) return 6;\n"
        '    printf("%s\\n", cab_status_string(CAB_ERR_NO_DEVICE));\n'
        "    return 0;\n}\n")
    libdir = os.path.join(root, "multimodal_audio_search_b200")
    exe = tmp_path / "main"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lcab", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stderr)
    assert "no CPU fallback" in out.stdout
