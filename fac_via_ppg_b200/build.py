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
    """Content hash of the sources and flags.  Paths enter relative to the package so that a library built in one
    checkout (the build container) is accepted as up to date in a copy of it (the GPU box)."""
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files.append(os.path.join(INCLUDE, "fac_b200.h"))
    for path in files:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _up_to_date(digest: str) -> bool:
    if not (os.path.isfile(LIB_PATH) and os.path.isfile(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as fh:
        return fh.read().strip() == digest


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into libfacb200.so; no-op if up to date.  Safe when several processes (the
    ranks of one torchrun launch) call it at once: one builds under a file lock into a temporary file that is
    renamed into place, the others wait and find the library up to date."""
    digest = _digest()
    if not force and _up_to_date(digest):
        return LIB_PATH
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(digest):       # another process built it while we waited
                return LIB_PATH
            nvcc = nvcc_path()
            if not os.path.isfile(nvcc):
                raise RuntimeError("nvcc not found; cannot build %s" % LIB_PATH)
            tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-o", tmp] + sources()
            proc = subprocess.run(cmd, capture_output=True, text=True)
            if proc.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stderr))
            if verbose:
                print(proc.stderr)
            os.replace(tmp, LIB_PATH)                    # atomic: a reader sees the old or the new file, never a torso
            with open(STAMP_PATH + ".tmp", "w") as fh:
                fh.write(digest)
            os.replace(STAMP_PATH + ".tmp", STAMP_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
