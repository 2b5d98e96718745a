"""Builds ``libgrapes_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Every ``csrc/*.cu`` is compiled to its own object (in parallel; cached under ``grapes_b200/build/`` by the sha1 of the
source, the headers and the flags) and the objects are linked into one shared library.  The library carries the sha1 of
its whole source tree (``grapes_build_id()``): the loader compares it with the tree it parses its ctypes prototypes
from, so a binary that does not belong to the header can never be called."""
from __future__ import annotations

import fcntl
import glob
import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, "libgrapes_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h")))


def _sha(paths, extra: str = "") -> str:
    h = hashlib.sha1(extra.encode())
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def tree_id() -> str:
    """sha1 over every source and header of the library as they are on disk now."""
    return _sha(sources() + _headers())


def built_id() -> str:
    """``grapes_build_id()`` of the library on disk ('' when it is missing or predates the build id)."""
    if not os.path.isfile(LIB_PATH):
        return ""
    import re
    with open(LIB_PATH, "rb") as f:               # read from the file: dlopen would pin the old image in this process
        m = re.search(rb"GRAPES_BUILD_ID=([0-9a-f]{40})", f.read())
    return m.group(1).decode() if m else ""


def is_stale() -> bool:
    return built_id() != tree_id()


def build_library(force: bool = False, verbose: bool = False) -> str:
    tid = tree_id()
    if not force and not verbose and built_id() == tid:
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found and libgrapes_b200.so is missing or does not match the source tree")
    os.makedirs(OBJ_DIR, exist_ok=True)
    # one builder at a time (torchrun starts one process per GPU; all of them may find a stale library)
    with open(os.path.join(OBJ_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not verbose and built_id() == tid:
            return LIB_PATH
        hdr = _sha(_headers(), " ".join(NVCC_FLAGS))
        extra = ["-Xptxas", "-v"] if verbose else []

        def compile_one(src: str):
            base = os.path.basename(src)[:-3]
            obj, keyf = os.path.join(OBJ_DIR, base + ".o"), os.path.join(OBJ_DIR, base + ".key")
            defs = [f'-DGRAPES_BUILD_ID="{tid}"'] if base == "ctx" else []
            key = _sha([src], hdr + "".join(defs))
            if not force and not verbose and os.path.isfile(obj) and os.path.isfile(keyf) and open(keyf).read() == key:
                return obj, ""
            res = subprocess.run([nvcc] + NVCC_FLAGS + extra + defs + ["-c", "-o", obj, src],
                                 capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {base}.cu:\n" + res.stdout + res.stderr)
            with open(keyf, "w") as f:
                f.write(key)
            return obj, res.stderr

        with ThreadPoolExecutor(max_workers=min(len(sources()), os.cpu_count() or 4)) as pool:
            out = list(pool.map(compile_one, sources()))
        tmp = LIB_PATH + ".tmp"
        res = subprocess.run([nvcc, "-shared", "-o", tmp] + [o for o, _ in out], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
        os.replace(tmp, LIB_PATH)
        if verbose:
            print("".join(log for _, log in out))
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
