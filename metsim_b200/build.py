"""Build ``csrc/libmetsim_b200.so`` in-tree with nvcc for sm_100a.

    python -m metsim_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
SOURCES = ("api.cu", "solar.cu", "daily.cu", "disagg.cu", "disagg_event.cu", "disagg_interior.cu",
           "disagg_streams.cu", "sink.cu")
HEADERS = ("msb_common.cuh", "disagg_common.cuh")
OUT = os.path.join(CSRC, "libmetsim_b200.so")


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


FLAGS = ("-O3", "-std=c++17", "--threads", "0", "-gencode", "arch=compute_100a,code=sm_100a",
         "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-shared")
STAMP = OUT + ".sha"      # hash of the sources the library was built from (travels with the .so)


def source_hash(extra: tuple = ()) -> str:
    """sha256 over every source, header and compiler flag that goes into the library (12 hex digits)."""
    h = hashlib.sha256()
    for path in [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.join(INCLUDE, "metsim_b200.h")]:
        h.update(os.path.basename(path).encode())
        h.update(open(path, "rb").read())
    h.update(" ".join(FLAGS + tuple(extra)).encode())
    return h.hexdigest()[:12]


def built_hash() -> str | None:
    """Hash recorded when the present .so was compiled (None: no library, or one of unknown origin)."""
    if os.path.exists(OUT) and os.path.exists(STAMP):
        return open(STAMP).read().strip()
    return None


def needs_build() -> bool:
    """By content, not by mtime: a shipped .so is reused only if it was built from exactly these
    sources and flags."""
    return built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False, bounds: bool = False) -> str:
    """``bounds=True`` compiles the debug checks (-DMSB_BOUNDS): every day-record / calendar /
    table access of the streaming kernels is range-checked on the device (trap or MSB_E_KNOTS)."""
    if not force and not needs_build():
        return OUT
    cmd = [_nvcc(), *FLAGS, "-I", INCLUDE, "-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    if os.path.exists(STAMP):
        os.remove(STAMP)
    if bounds:
        cmd.insert(1, "-DMSB_BOUNDS")
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    with open(STAMP, "w") as fh:          # a debug build never passes for the product library
        fh.write(source_hash(("-DMSB_BOUNDS",) if bounds else ()) + "\n")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv or "--bounds" in sys.argv, verbose="--verbose" in sys.argv,
                bounds="--bounds" in sys.argv))
