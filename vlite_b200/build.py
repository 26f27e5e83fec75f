"""Build libvlite_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m vlite_b200.build [--force] [--verbose]

The library is plain CUDA runtime + C ABI (no torch, no Python in the link), so
any host language can bind it.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libvlite_b200.so")
SOURCES = [os.path.join(CSRC, "api.cu")]
HEADERS = [os.path.join(CSRC, f) for f in ("common.cuh", "scan.cuh", "mmascan.cuh", "merge.cuh", "rescore.cuh",
                                            "repack.cuh", "synth.cuh", "exchange.cuh", "largek.cuh", "host_util.cuh",
                                            "corpus.cuh", "search_host.cuh", "group.cuh")]
HEADERS.append(os.path.join(os.path.dirname(PKG), "include", "vlite_b200.h"))


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS if os.path.exists(f))


def build(force: bool = False, verbose: bool = False, out: str | None = None, extra_flags=()) -> str:
    """out / extra_flags: experiment builds (e.g. -DVL_RMAX=2) next to the shipped library."""
    if out is None and not force and not is_stale():
        return LIB_PATH
    out = out or LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [
        _nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "--shared", "-Xcompiler", "-fPIC,-O3,-Wall,-pthread",
        "-o", out,
    ] + list(extra_flags) + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libvlite_b200.so")
    return out


DEBUG_LIB_PATH = os.path.join(LIB_DIR, "libvlite_b200_dbg.so")


def build_debug() -> str:
    """the bounds-checked twin (-DVL_DEBUG_BOUNDS); load it with VLITE_B200_LIB=<path>"""
    return build(out=DEBUG_LIB_PATH, extra_flags=["-DVL_DEBUG_BOUNDS"])


if __name__ == "__main__":
    if "--debug" in sys.argv:
        print(build_debug())
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
