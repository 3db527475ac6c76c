"""Build recipe for libpsgb200.so (nvcc, sm_100a only, in-tree output).

``python -m pyspectrogram_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the .so is git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "psg_b200.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "sti_kernels.cuh"), os.path.join(HERE, "csrc", "sti_cluster.cuh"), os.path.join(HERE, "csrc", "sti_whole.cuh"), os.path.join(HERE, "csrc", "sti_whole16.cuh"), os.path.join(HERE, "csrc", "sti_bluestein.cuh"), os.path.join(HERE, "csrc", "cplx.cuh"),
        os.path.join(HERE, "..", "include", "psg_b200.h")]
OUT = os.path.join(HERE, "libpsgb200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUT
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", OUT, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
