"""Build real2sim_eval_b200/libr2s.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m real2sim_eval_b200.build [--force] [--verbose]

The library is plain CUDA C++ behind an extern "C" boundary (include/*.h): no
torch headers, static cudart, so it loads in any process (also on a box without
a GPU, where only symbol checks run).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libr2s.so")
SOURCES = ["common.cu", "phys.cu", "raster.cu", "lbs.cu", "links.cu", "eef.cu", "metrics.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libr2s.so")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    host_cc = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None
    common = [_nvcc(), *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if host_cc:
        common += ["-ccbin", host_cc]
    if verbose:
        common += ["-Xptxas", "-v"]
    common += os.environ.get("R2S_NVCC_FLAGS", "").split()
    out = os.environ.get("R2S_LIB_OUT", OUT)
    procs = []
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append(subprocess.Popen(common + ["-c", os.path.join(CSRC, src), "-o", obj]))
    for pr in procs:
        if pr.wait() != 0:
            raise RuntimeError("nvcc failed building libr2s.so")
    subprocess.run(common + ["-shared", *objs, "-o", out], check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
