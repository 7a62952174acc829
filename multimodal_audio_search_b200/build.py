"""In-tree build of the CUDA C-ABI library (libcab.so) for sm_100a with nvcc.

    python -m multimodal_audio_search_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box
with the repo snapshot.  Objects are rebuilt only when a source/header is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libcab.so")
TORCH_LIB = os.path.join(HERE, "libcab_torch.so")
TORCH_SRC = os.path.join(CSRC, "torch_binding.cpp")

SOURCES = ["cab_api.cu", "cab_ingest.cu", "cab_gemv.cu", "cab_finalize.cu",
           "cab_gemm_tc.cu", "cab_score_all.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v",
              "--expt-relaxed-constexpr"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def _deps_mtime() -> float:
    m = 0.0
    for d in (CSRC, INCLUDE):
        for f in os.listdir(d):
            if f.endswith((".h", ".cuh")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return max(m, os.path.getmtime(__file__))


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    hdr = _deps_mtime()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(hdr, os.path.getmtime(s)):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        with open(o + ".log", "w") as f:
            f.write(r.stdout + r.stderr)
        return s, r.stdout + r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for s, log in ex.map(compile_one, jobs):
                if verbose:
                    print(f"== {os.path.basename(s)}\n{log}")
    objs = [os.path.join(BUILD, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs,
               "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_torch_extension(force: bool = False) -> str:
    """PyTorch extension over the C-ABI (csrc/torch_binding.cpp -> libcab_torch.so): registers
    torch.ops.cab.*; links libcab.so via $ORIGIN.  Plain g++ against the installed torch headers."""
    import torch
    from torch.utils import cpp_extension as ce
    newest = max(os.path.getmtime(TORCH_SRC), os.path.getmtime(os.path.join(INCLUDE, "cab.h")), os.path.getmtime(LIB))
    if not force and os.path.exists(TORCH_LIB) and os.path.getmtime(TORCH_LIB) >= newest:
        return TORCH_LIB
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", TORCH_SRC, "-o", TORCH_LIB]
    for inc in ce.include_paths() + [cuda_inc]:
        cmd += ["-isystem", inc]
    for lp in ce.library_paths():
        cmd += [f"-L{lp}", f"-Wl,-rpath,{lp}"]
    cmd += [f"-L{HERE}", "-Wl,-rpath,$ORIGIN", "-lcab", "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-ltorch_cuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"torch extension build failed:\n{r.stdout}\n{r.stderr}")
    return TORCH_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_torch_extension(force="--force" in sys.argv))
