"""Builds libfacb200.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

The shared object lives next to this file so that it travels to the GPU box with
the repository snapshot; it is git-ignored.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libfacb200.so")
STAMP_PATH = os.path.join(HERE, ".libfacb200.stamp")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files.append(os.path.join(INCLUDE, "fac_b200.h"))
    for path in files:
        h.update(path.encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into libfacb200.so; no-op if up to date."""
    digest = _digest()
    if not force and os.path.isfile(LIB_PATH) and os.path.isfile(STAMP_PATH):
        with open(STAMP_PATH) as fh:
            if fh.read().strip() == digest:
                return LIB_PATH
    nvcc = nvcc_path()
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found; cannot build %s" % LIB_PATH)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-o", LIB_PATH] + sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stderr))
    if verbose:
        print(proc.stderr)
    with open(STAMP_PATH, "w") as fh:
        fh.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
