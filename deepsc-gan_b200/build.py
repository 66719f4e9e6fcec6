"""Build libdeepsc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Every ``csrc/*.cu`` is compiled to an object under ``csrc/_obj/`` (in parallel, skipped when the object is newer than the
source and every header), then linked into ``csrc/libdeepsc_b200.so``.  ``DSC_DEBUG_TOOLS=1`` additionally compiles the
sources under ``csrc/debug/`` (micro-benchmarks and the star-kernel timeline, not part of the product ABI) into
``csrc/libdeepsc_b200_debug.so``.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
OUT = os.path.join(CSRC, "libdeepsc_b200.so")
OUT_DEBUG = os.path.join(CSRC, "libdeepsc_b200_debug.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I", CSRC]


def sources(debug: bool = False):
    """Product sources; the debug-tools library is the same sources compiled with -DDSC_DEBUG_TOOLS=1 plus csrc/debug/*.cu."""
    src = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    return src + sorted(glob.glob(os.path.join(CSRC, "debug", "*.cu"))) if debug else src


def _headers():
    return (glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "debug", "*.cuh"))
            + glob.glob(os.path.join(HERE, "..", "include", "*.h")))


def needs_build(debug: bool = False) -> bool:
    out = OUT_DEBUG if debug else OUT
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in sources(debug) + _headers())


def _compile(nvcc, src, obj, verbose, extra):
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """Returns the path of the shared library (the debug-tools library when ``debug``)."""
    out = OUT_DEBUG if debug else OUT
    if not force and not needs_build(debug):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    jobs, objs = [], []
    for src in sources(debug):
        obj = os.path.join(OBJ, ("dbg_" if debug else "") + os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or verbose or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))
    extra = ["-DDSC_DEBUG_TOOLS=1"] if debug else []
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda so: _compile(nvcc, so[0], so[1], verbose, extra), jobs))
    failed = [(s, r) for s, r in results if r.returncode != 0]
    for s, r in results:
        if r.returncode != 0 or verbose:
            sys.stderr.write(f"---- {os.path.basename(s)}\n{r.stdout}{r.stderr}")
    if failed:
        raise RuntimeError("nvcc failed on " + ", ".join(os.path.basename(s) for s, _ in failed))
    r = subprocess.run([nvcc, "-shared", "-o", out] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"link of {os.path.basename(out)} failed")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
