"""Builds metalchat_b200/libmc_cuda.so in-tree with nvcc for sm_100a.

``python -m metalchat_b200.build`` (or ``__graft_entry__.build()``).  nvcc cross-compiles
without a GPU.  Objects are cached under metalchat_b200/csrc/_obj and rebuilt when a source
or header is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = CSRC / "_obj"
LIB = HERE / "libmc_cuda.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libmc_cuda.so cannot be built (there is no CPU fallback)")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _headers() -> list[Path]:
    return sorted(CSRC.glob("*.cuh")) + sorted((HERE.parent / "include").glob("*.h"))


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    OBJ.mkdir(parents=True, exist_ok=True)
    hdr_mtime = max(h.stat().st_mtime for h in _headers())
    jobs = []
    objs = []
    for src in sources():
        obj = OBJ / (src.stem + ".o")
        objs.append(obj)
        stale = force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_mtime)
        if stale:
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("MC_NVCC_EXTRA", "").split(), "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (src.stem + ".ptxas.log")).write_text(res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stderr[-6000:]}")
        if verbose:
            print(res.stderr, file=sys.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not LIB.exists() or force:
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-lcudart", "-ldl", "-Xlinker", "-z,defs"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stderr[-4000:]}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
