# -*- coding: utf-8 -*-
"""Builds cellvit_b200/libcellvit_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree, no JIT cache).

Run ``python -m cellvit_b200.build``; ``__graft_entry__.build()`` calls :func:`build`. nvcc cross-compiles
without a GPU. Objects go to build/obj (git-ignored); the .so sits next to this file so that it
travels to the GPU box with the source snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(os.path.dirname(HERE), "build", "obj")
LIB = os.path.join(HERE, "libcellvit_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _headers_digest() -> str:
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".h", ".cuh")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, digest: str) -> str:
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    stamp = obj + ".stamp"
    key = digest + hashlib.sha1(open(src, "rb").read()).hexdigest()
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == key:
        return obj
    r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
    log = r.stdout + r.stderr
    open(obj + ".log", "w").write(log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{log}")
    open(stamp, "w").write(key)
    return obj


def build(verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    digest = _headers_digest()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, digest), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    if verbose:
        for o in objs:
            for line in open(o + ".log").read().splitlines():
                if "registers" in line or "spill" in line and "0 bytes spill stores" not in line:
                    print(os.path.basename(o), line.strip())
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
