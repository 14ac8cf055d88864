"""Build recipe for libnwwb200.so — explicit nvcc for sm_100a, in-tree output.

Run ``python -m nanowakeword_b200.build`` (or ``__graft_entry__.build()``).  The shared
library lands next to the sources (``nanowakeword_b200/csrc/libnwwb200.so``) so that it
travels with the tree; it is git-ignored.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(CSRC, "libnwwb200.so")
SOURCES = ["nww_engine.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libnwwb200.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(CSRC), "..", "include", "nww_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, *SOURCES, "-o", LIB_PATH]
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if verbose:
        for line in (proc.stdout + proc.stderr).splitlines():
            if "registers" in line or "error" in line or "warning" in line or "spill" in line:
                print(line)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libnwwb200.so")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
