"""
In-tree build of ``libfnsm_b200.so`` (sm_100a only) with nvcc.

``python -m feinsum_b200._build`` or ``__graft_entry__.build()``.  Objects are
cached under ``feinsum_b200/csrc/_obj`` keyed by a hash of the source, the
headers and the flags, compiled in parallel, then linked into
``feinsum_b200/libfnsm_b200.so`` -- the file the ctypes loader opens and that
travels to the GPU box.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(CSRC, "_obj")
LIB_PATH = os.path.join(PKG_DIR, "libfnsm_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libfnsm_b200.so")


def _digest(paths: list[str], extra: str) -> str:
    h = hashlib.sha1(extra.encode())
    for p in paths:
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def sources() -> list[str]:
    return sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu")
    )


def headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hs += [os.path.join(INCLUDE_DIR, f) for f in os.listdir(INCLUDE_DIR) if f.endswith(".h")]
    return sorted(hs)


def build(verbose: bool = False, force: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = headers()
    flags = [*NVCC_FLAGS, "-I", INCLUDE_DIR]
    if verbose:
        flags += ["-Xptxas", "-v"]
    jobs = []
    objs = []
    for src in sources():
        key = _digest([src, *hdrs], " ".join(flags))
        obj = os.path.join(OBJ_DIR, f"{os.path.basename(src)[:-3]}.{key}.o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            jobs.append((src, obj))

    def compile_one(job: tuple[str, str]) -> None:
        src, obj = job
        cmd = [nvcc, *flags, "-c", src, "-o", obj + ".tmp.o"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}")
        os.replace(obj + ".tmp.o", obj)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            list(pool.map(compile_one, jobs))
    # drop stale objects
    keep = {os.path.basename(o) for o in objs}
    for f in os.listdir(OBJ_DIR):
        if f.endswith(".o") and f not in keep:
            os.remove(os.path.join(OBJ_DIR, f))
    link_key = _digest(objs, "link")
    stamp = os.path.join(OBJ_DIR, "link.stamp")
    old = open(stamp).read() if os.path.exists(stamp) else ""
    if force or jobs or old != link_key or not os.path.exists(LIB_PATH):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
               "-o", LIB_PATH + ".tmp", *objs, "-lcudart"]
        subprocess.run(cmd, check=True)
        os.replace(LIB_PATH + ".tmp", LIB_PATH)
        with open(stamp, "w") as fh:
            fh.write(link_key)
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
