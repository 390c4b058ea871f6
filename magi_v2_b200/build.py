"""Build libmagi_b200.so (the C-ABI library, include/magi_b200.h) in-tree with nvcc for sm_100a.

    python -m magi_v2_b200.build [--force]

Incremental: a translation unit is recompiled only when it or a header changed.  The library
has no dependency on torch or Python; the built .so sits next to this file so that it travels
to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmagi_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "magi_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


# headers only one translation unit includes (kept out of _headers_mtime so that touching them does not rebuild
# sampler.cu, which takes minutes)
EXTRA_DEPS = {"nuts.cu": ["magi_b200_nuts.h"], "posterior_wide.cu": ["magi_b200_wide.h"], "probe.cu": ["magi_b200_probe.h"]}


def _compile(src, force, hm):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    sp = os.path.join(CSRC, src)
    newest = max([os.path.getmtime(sp), hm] +
                 [os.path.getmtime(os.path.join(HERE, "..", "include", h)) for h in EXTRA_DEPS.get(src, [])])
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > newest:
        return obj, False
    cmd = [NVCC] + FLAGS + ["-c", sp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, True


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hm = _headers_mtime()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force, hm), srcs))
    objs = [o for o, _ in res]
    changed = any(c for _, c in res)
    if changed or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                     "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[magi_v2_b200.build] {'rebuilt' if changed else 'up to date'}: {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
