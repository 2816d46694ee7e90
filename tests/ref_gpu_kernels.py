"""Launch the REFERENCE's own CUDA kernels (precompiled into baseline/_ref/*.cubin by
tools/make_ref_gpu_baseline.py) through cuda-python, exactly as the reference's wrappers do
(fields.py:823-852 for forces, :980-1013 for the potential): grid ceil(N/128), block 128,
float4 (x,y,z,m) packing for float32, SoA doubles for float64, eps2 = dtype(1e-15).

Test / measurement infrastructure: lets the GPU tests compare this package with the outputs of the
reference's GPU code on the same B200, and tools/ref_gpu_bench.py time it as a second baseline."""
from __future__ import annotations

import ctypes
import json
from pathlib import Path

import numpy as np

REF_DIR = Path(__file__).resolve().parents[1] / "baseline" / "_ref"
_modules: dict = {}


def available() -> bool:
    return (REF_DIR / "manifest.json").exists()


def _function(key: str):
    from cuda.bindings import driver as cu

    if key not in _modules:
        man = json.loads((REF_DIR / "manifest.json").read_text())[key]
        data = (REF_DIR / man["cubin"]).read_bytes()
        err, mod = cu.cuModuleLoadData(data)
        assert err == cu.CUresult.CUDA_SUCCESS, err
        err, fn = cu.cuModuleGetFunction(mod, man["kernel"].encode())
        assert err == cu.CUresult.CUDA_SUCCESS, err
        _modules[key] = (mod, fn, man)
    return _modules[key][1], _modules[key][2]


def launch(kind: str, precision: str, pos_t, mass_t, h_t, G: float, kernel_id: int):
    """kind 'force' | 'potential'; pos_t (N,3), mass_t (N,), h_t (N,) torch CUDA tensors of the kernel dtype.
    Returns the output tensor ((N,3) or (N,)); asynchronous on torch's current stream."""
    import torch
    from cuda.bindings import driver as cu

    key = f"{kind}_{precision}"
    fn, man = _function(key)
    N = pos_t.shape[0]
    dt = pos_t.dtype
    stream = torch.cuda.current_stream().cuda_stream
    blocks = (N + 127) // 128
    fl = ctypes.c_float if dt == torch.float32 else ctypes.c_double
    n_out = 3 if kind == "force" else 1
    outs = [torch.empty(N, dtype=dt, device=pos_t.device) for _ in range(n_out)]
    if man["float4"]:
        pm = torch.empty((N, 4), dtype=dt, device=pos_t.device)      # fields.py:832-836
        pm[:, :3] = pos_t
        pm[:, 3] = mass_t
        ptrs = [pm.data_ptr(), h_t.data_ptr()] + [o.data_ptr() for o in outs]
        keep = [pm]
    else:
        cols = [pos_t[:, i].contiguous() for i in range(3)]           # fields.py:596-598
        ptrs = [c.data_ptr() for c in cols] + [mass_t.data_ptr(), h_t.data_ptr()] + [o.data_ptr() for o in outs]
        keep = cols
    vals = tuple(ptrs) + (G, 1e-15, kernel_id, N)
    types = tuple([ctypes.c_void_p] * len(ptrs)) + (fl, fl, ctypes.c_int, ctypes.c_int)
    (err,) = cu.cuLaunchKernel(fn, blocks, 1, 1, 128, 1, 1, 0, stream, (vals, types), 0)
    assert err == cu.CUresult.CUDA_SUCCESS, err
    out = torch.stack(outs, dim=1) if kind == "force" else outs[0]
    out._keepalive = keep
    return out
