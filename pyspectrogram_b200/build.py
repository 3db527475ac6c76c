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
CSRC = os.path.join(HERE, "csrc")
# translation units (compiled in parallel, linked into one library); every header is a dependency of both
UNITS = ["psg_b200.cu", "psg_r32.cu", "psg_mixct.cu"]
DEPS = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h", ".inc"))] + [
    os.path.join(HERE, "..", "include", "psg_b200.h")]
OUT = os.path.join(HERE, "libpsgb200.so")
OBJ_DIR = os.path.join(HERE, "build")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMPILE_FLAGS = [*ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
LINK_FLAGS = [*ARCH_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]


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


# headers only the small units include (editing them does not rebuild the big unit), and what each small unit sees
SMALL_ONLY = ("sti_r32.cuh", "r32_math.cuh", "sti_mixct.cuh", "mixct_plans.inc")
UNIT_HDRS = {
    "psg_r32.cu": ("psg_r32.h", "sti_r32.cuh", "r32_math.cuh", "sti_common.cuh", "cplx.cuh"),
    "psg_mixct.cu": ("psg_mixct.h", "sti_mixct.cuh", "mixct_plans.inc", "sti_common.cuh", "cplx.cuh"),
}


def _unit_stale(src: str, obj: str) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    hdrs = [d for d in DEPS if not d.endswith(".cu")]
    if os.path.basename(src) in UNIT_HDRS:
        hdrs = [d for d in hdrs if os.path.basename(d) in UNIT_HDRS[os.path.basename(src)]]
    else:
        hdrs = [d for d in hdrs if os.path.basename(d) not in SMALL_ONLY]
    return any(os.path.getmtime(d) > t for d in [src, *hdrs] if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUT
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    procs = []
    objs = []
    for unit in UNITS:
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJ_DIR, unit[:-3] + ".o")
        objs.append(obj)
        if not force and not _unit_stale(src, obj):
            continue
        cmd = [nvcc, *COMPILE_FLAGS, "-c", "-o", obj, src]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + out)
        if verbose:
            print(out)
    # link next to the target and rename: a reader (a test run, a snapshot of the tree) never sees a half-written library
    tmp = OUT + ".tmp"
    cmd = [nvcc, *LINK_FLAGS, "-o", tmp, *objs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
