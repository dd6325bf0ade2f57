"""Build libpmr_b200.so (sm_100a) in-tree with nvcc.  No JIT cache, no torch extension: the
shared library is a plain C-ABI object that travels with the repo snapshot to the GPU box."""

from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
REPO = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "_lib"
LIB = LIBDIR / "libpmr_b200.so"

SOURCES = ["pmr_gemm.cu", "pmr_factor.cu", "pmr_assemble.cu", "pmr_ao2mo.cu", "pmr_misc.cu", "pmr_eri.cu", "pmr_eig.cu", "pmr_iter.cu"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libpmr_b200.so cannot be built")


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / "pmr_common.cuh", REPO / "include" / "pmr_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    LIBDIR.mkdir(exist_ok=True)
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)
    common = [
        nvcc_path(), "-std=c++17", "-O3", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-Xcompiler", "-fPIC", "-I", str(REPO / "include"), "-I", str(CSRC),
    ]
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    objs = []
    for s in SOURCES:
        obj = objdir / (s + ".o")
        objs.append(str(obj))
        src = CSRC / s
        if not force and obj.exists() and obj.stat().st_mtime > max(
            src.stat().st_mtime,
            (CSRC / "pmr_common.cuh").stat().st_mtime,
            (REPO / "include" / "pmr_b200.h").stat().st_mtime,
        ):
            continue
        procs.append((s, subprocess.Popen(common + ["-c", str(src), "-o", str(obj)],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
    link = [nvcc_path(), "-shared", "-o", str(LIB)] + objs + ["-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
