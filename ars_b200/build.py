"""Build libars_b200.so (hand-written sm_100a CUDA kernels + the C ABI) in-tree with nvcc.

    python -m ars_b200.build [--force]

The shared object lands in ars_b200/lib/ (git-ignored, but shipped to the GPU box).
nvcc cross-compiles without a GPU, so this runs in the authoring container too.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libars_b200.so")
SOURCES = ["fft_k_gc_fwd.cu", "fft_k_gc_inv.cu", "fft_k_gs_fwd.cu", "fft_k_gs_inv.cu", "fft_k_fs_fwd.cu", "fft_k_fs_inv.cu",
           "fft_k_fc_fwd.cu", "fft_k_fc_inv.cu", "fft_k_mid.cu", "common.cu", "hostio.cu", "fft_plan.cu", "spectral.cu", "epilogue.cu", "ir_synth.cu", "metrics.cu", "upols.cu",
           "api.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Wno-deprecated-gpu-targets"]
if os.environ.get("ARS_EXTRA_NVCC_FLAGS"):          # experiments, e.g. -DARS_RAD13=1
    NVCC_FLAGS += os.environ["ARS_EXTRA_NVCC_FLAGS"].split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libars_b200 cannot be built (there is no CPU implementation)")


def _deps_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(os.path.dirname(HERE), "include", "ars_b200.h"))
    return max(os.path.getmtime(p) for p in paths)


def _compile(src: str) -> str:
    obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
    srcp = os.path.join(CSRC, src)
    if os.path.exists(obj) and os.path.getmtime(obj) >= _deps_mtime():
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if force:
        for f in os.listdir(BUILD):
            os.remove(os.path.join(BUILD, f))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-Wno-deprecated-gpu-targets"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {LIB} ({os.path.getsize(LIB) / 1e6:.1f} MB)")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
