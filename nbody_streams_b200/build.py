"""Build libnbody_sm100.so (hand-written sm_100a CUDA) in-tree with nvcc.

    python -m nbody_streams_b200.build [--force] [--jobs N]

The shared library is written next to this file so that it travels with the
repository snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
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
OBJ = HERE / "build"
LIB = HERE / "libnbody_sm100.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
] + os.environ.get("NB_EXTRA_NVCC_FLAGS", "").split()     # experiments only, e.g. -DNB_ACC_SCALAR


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA toolkit is required to build libnbody_sm100.so")
    return nvcc


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, jobs: int | None = None, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link the shared library."""
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "nbody_sm100.h"]
    srcs = _sources()
    todo = []
    for s in srcs:
        o = OBJ / (s.stem + ".o")
        if force or _stale(o, [s] + headers):
            todo.append((s, o))

    def compile_one(so):
        s, o = so
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (s.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s.name}:\n{r.stderr[-4000:]}")
        return s.name

    if todo:
        with ThreadPoolExecutor(max_workers=jobs or min(8, os.cpu_count() or 4)) as ex:
            for name in ex.map(compile_one, todo):
                if verbose:
                    print(f"  compiled {name}", flush=True)
    objs = [OBJ / (s.stem + ".o") for s in srcs]
    if force or todo or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs),
               "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
        if verbose:
            print(f"  linked {LIB.name}", flush=True)
    return LIB


if __name__ == "__main__":
    jobs = None
    if "--jobs" in sys.argv:
        jobs = int(sys.argv[sys.argv.index("--jobs") + 1])
    print(build(force="--force" in sys.argv, jobs=jobs, verbose=True))
