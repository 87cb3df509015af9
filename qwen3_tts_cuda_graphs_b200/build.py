"""Build the in-tree CUDA libraries with nvcc for sm_100a (explicit -gencode; nothing is JIT-compiled).

`python -m qwen3_tts_cuda_graphs_b200.build` or `__graft_entry__.build()`.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
REPO = os.path.dirname(PKG_DIR)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr", "-diag-suppress", "550",
]

LIBS = {
    "libfq3.so": (["fq3_api.cu"], ["fq3_kernel.cuh", "fq3_common.cuh", "../../include/fq3.h"], []),
    "libfq3codec.so": (["fq3_codec.cu"], ["../../include/fq3_codec.h"], []),
}


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA libraries must be prebuilt in-tree (run build() where nvcc exists)")
    return exe


def lib_path(name: str) -> str:
    return os.path.join(PKG_DIR, name)


def is_stale(name: str) -> bool:
    srcs, deps, _ = LIBS[name]
    out = lib_path(name)
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    for f in srcs + deps:
        fp = os.path.join(CSRC, f)
        if os.path.exists(fp) and os.path.getmtime(fp) > t:
            return True
    return False


def build_lib(name: str, force: bool = False, verbose: bool = False) -> str:
    srcs, _, extra = LIBS[name]
    out = lib_path(name)
    if not all(os.path.exists(os.path.join(CSRC, s)) for s in srcs):
        return out
    if not force and not is_stale(name):
        return out
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-o", out] + [os.path.join(CSRC, s) for s in srcs]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return out


def build_all(force: bool = False, verbose: bool = False):
    return [build_lib(n, force=force, verbose=verbose) for n in LIBS]


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", [n for n in LIBS if os.path.exists(lib_path(n))])
