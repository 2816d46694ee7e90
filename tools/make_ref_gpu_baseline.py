#!/usr/bin/env python3
"""Compile the REFERENCE's own CUDA kernels for the B200 into baseline/_ref/ (git-ignored).

Runs in the build container only (needs /root/reference).  The reference JIT-compiles its CUDA-C template
strings with CuPy at run time (fields.py:417-483: `template.format(**_TYPE_SPECS[precision])`, nvcc options
`-O3 --use_fast_math -arch=sm_{cc}`); CuPy is not available here or on the GPU box, so the SAME formatting
and the SAME nvcc flags are applied ahead of time (arch = sm_100) and the cubins travel with the snapshot.
tools/ref_gpu_bench.py / tests/test_gpu_vs_reference_kernels.py load them through cuda-python and launch
them exactly as fields.py:823-852 / :980-1013 do (grid ceil(N/128), block 128).

This gives (a) "the reference GPU kernel on the same B200" as a second baseline and (b) outputs of the
reference's own GPU code to check parity against.  No reference source is committed: the generated .cu and
.cubin files live only under baseline/_ref/.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import types
import warnings
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parents[1] / "baseline" / "_ref"


def main() -> int:
    if not REF.exists():
        print("no /root/reference: nothing to do")
        return 0
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nb_numba_cache")
    sys.dont_write_bytecode = True
    warnings.simplefilter("ignore")
    pkg = types.ModuleType("nbody_streams")
    pkg.__path__ = [str(REF / "nbody_streams")]
    sys.modules["nbody_streams"] = pkg
    import importlib

    fields = importlib.import_module("nbody_streams.fields")
    OUT.mkdir(parents=True, exist_ok=True)
    manifest = {}
    nvcc = "/usr/local/cuda/bin/nvcc"
    for (kind, precision, kahan), (template, name) in fields._NBODY_KERNEL_CONFIG.items():
        if precision == "float64" and kahan:
            continue   # unreachable from the public API (fields.py:820)
        key = f"{kind}_{precision}{'_kahan' if kahan else ''}"
        src = template.format(**fields._TYPE_SPECS[precision])
        cu = OUT / f"{key}.cu"
        cu.write_text(src)
        cubin = OUT / f"{key}.cubin"
        cmd = [nvcc, "-O3", "--use_fast_math", "-arch=sm_100", "-cubin", "-o", str(cubin), str(cu)]
        subprocess.run(cmd, check=True, capture_output=True)
        cu.unlink()   # keep only the binary: no reference source lingers in the tree
        manifest[key] = {"kernel": name, "cubin": cubin.name, "kind": kind, "precision": precision, "kahan": kahan,
                         "float4": precision == "float32"}
        print(f"{key:28s} -> {cubin.name} ({cubin.stat().st_size} B)")
    (OUT / "manifest.json").write_text(json.dumps(manifest, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
