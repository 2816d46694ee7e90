"""The C ABI used from plain C (no Python objects, no torch): builds tests/capi/capi_smoke.c against
include/nbody_sm100.h + libnbody_sm100.so and runs it on the GPU."""
import os
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "capi" / "capi_smoke.c"
LIBDIR = ROOT / "nbody_streams_b200"


def _build(tmp_path):
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    exe = tmp_path / "capi_smoke"
    cmd = [gcc, "-O2", "-std=gnu11", str(SRC), "-I", str(ROOT / "include"), "-I/usr/local/cuda/include",
           "-L", str(LIBDIR), "-lnbody_sm100", "-L/usr/local/cuda/lib64", "-lcudart", "-lm",
           f"-Wl,-rpath,{LIBDIR}", "-Wl,-rpath,/usr/local/cuda/lib64", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_program_compiles_and_links_against_the_header(tmp_path):
    """CPU side: the header is valid C and every symbol the program uses resolves against the library."""
    assert _build(tmp_path).exists()


@pytest.mark.gpu
def test_c_program_runs_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([str(exe), "3001"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "CAPI SMOKE OK" in r.stdout, r.stdout + r.stderr
