"""In-tree build of the sm_100a CUDA library (``lib/libprv2_b200.so``) with nvcc.

nvcc cross-compiles without a GPU, so this runs on the CPU-only build box; the resulting .so is
git-ignored but travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libprv2_b200.so")
SOURCES = ["geometry.cu", "pointwise.cu", "umma_gemm.cu", "attention.cu", "zoe_head.cu"]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "prv2_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
] + os.environ.get("PRV2_EXTRA_NVCC_FLAGS", "").split()        # diagnostics builds, e.g. -DPRV2_GEMM_TRACE_BUILD -DPRV2_ATTN_TRACE_BUILD


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libprv2_b200.so")


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def built_digest() -> str:
    """Digest stamped into the existing .so, '' when there is none / it predates the stamp.  Read from the file's bytes
    (marker ``PRV2_DIGEST=``), not through dlopen: a dlopen'ed handle would survive the rebuild that may follow."""
    if not os.path.exists(LIB_PATH):
        return ""
    with open(LIB_PATH, "rb") as fh:
        blob = fh.read()
    i = blob.find(b"PRV2_DIGEST=")
    return blob[i + 12:i + 12 + 64].decode("ascii", "replace") if i >= 0 else ""


def source_digest() -> str:
    return _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    dig = _digest()
    if not force and os.path.exists(LIB_PATH) and built_digest() == dig:
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, f'-DPRV2_BUILD_DIGEST="{dig}"', "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
