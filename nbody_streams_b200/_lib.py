"""ctypes binding of libnbody_sm100.so (C ABI in include/nbody_sm100.h).

The reference's precedent for "Python -> thin C ABI -> CUDA .so" is its tree
code loader (reference nbody_streams/tree_gpu/_force.py:96-137); the direct
path there goes through cp.RawKernel instead (fields.py:479, :841-852), which
this library replaces.

There is NO CPU fallback: if the shared library or a CUDA device is missing the
callers raise ImportError (the reference raises ImportError("CuPy required"),
fields.py:786-787).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libnbody_sm100.so"

# enums (mirror include/nbody_sm100.h)
PREC_F32, PREC_F32_COMP, PREC_F64 = 0, 1, 2
MODE_FORCE, MODE_POT, MODE_BOTH = 1, 2, 3
DT_F32, DT_F64 = 0, 1

_lib = None

_vp, _i, _l, _d, _sz = C.c_void_p, C.c_int, C.c_long, C.c_double, C.c_size_t

# name -> (restype, argtypes); every symbol include/nbody_sm100.h declares
SIGNATURES = {
    "nb_version": (_i, []),
    "nb_error_string": (C.c_char_p, [_i]),
    "nb_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz), C.POINTER(_sz),
                            C.POINTER(_i), C.c_char_p, _i]),
    "nb_tile_j": (_i, []),
    "nb_block_i": (_i, [_i]),
    "nb_packed_bytes": (_sz, [_i, _l]),
    "nb_padded_count": (_l, [_l]),
    "nb_pack": (_i, [_i, _i, _vp, _l, _vp, _i, _vp, _i, _d, _vp, _i, _d, _l, _l, _d, _vp]),
    "nb_sweep_grid": (_i, [_i, _i, _i, C.POINTER(_i)]),
    "nb_workspace_bytes": (_sz, [_i, _l, _i, _i]),
    "nb_sweep": (_i, [_i, _i, _i, _vp, _l, _i, _i, _i, _i, _l, _l, _d, _vp, _sz, _i, _i, _vp]),
    "nb_finish": (_i, [_i, _i, _i, _l, _i, _i, C.POINTER(_i), _vp, _d, _vp, _vp, _vp]),
    "nb_direct": (_i, [_i, _i, _i, _vp, _i, _vp, _i, _d, _vp, _i, _d, _l, _d, _d, _vp, _vp, _sz, _i, _vp, _vp,
                       _vp]),
    "nb_sym_min_n": (_l, []),
    "nb_set_sym_min_n": (_i, [_l]),
    "nb_sym_block_tiles": (_i, [_i]),
    "nb_sym_grid": (_i, [_i, _i, _i, C.POINTER(_i)]),
    "nb_sym_workspace_bytes": (_sz, [_l]),
    "nb_sym_sweep": (_i, [_i, _i, _i, _vp, _l, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _d, _vp, _sz, _l, _i, _i,
                          _vp]),
    "nb_sym_finish": (_i, [_i, _i, _vp, _l, _l, _d, _vp, _vp, _vp]),
    "nb_sym_finish_peers": (_i, [_i, _i, C.POINTER(_vp), _i, _l, _l, _d, _vp, _vp, _vp]),
    "nb_direct_sym": (_i, [_i, _i, _i, _vp, _i, _vp, _i, _d, _vp, _i, _d, _l, _d, _d, _vp, _vp, _sz, _i, _vp, _vp,
                           _vp]),
    "nb_kick_drift_pack": (_i, [_i, _vp, _vp, _vp, _i, _vp, _vp, _l, _l, _d, _d, _vp]),
    "nb_kick_drift_push": (_i, [_i, _vp, _vp, _vp, _i, _vp, C.POINTER(_vp), _i, _vp, _l, _l, _d, _d, _vp]),
    "nb_kick": (_i, [_vp, _vp, _i, _vp, _l, _d, _vp]),
    "nb_energy_workspace_bytes": (_sz, [_l]),
    "nb_energy": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _l, _vp, _vp, _vp]),
    "nb_set_variant": (_i, [_i, _i]),
    "nb_get_variant": (_i, [_i]),
    "nb_ubench": (_i, [_i, _i, C.POINTER(_d), C.POINTER(_d)]),
}


class NbodyLibError(RuntimeError):
    """Non-zero return code from libnbody_sm100 (CUDA error or invalid argument)."""


def load() -> C.CDLL:
    """Load (once) and return the shared library; ImportError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("NBODY_SM100_LIB", LIB_PATH))
    if not path.exists():
        raise ImportError(
            f"libnbody_sm100.so not found at {path}. Build it with "
            "`python -m nbody_streams_b200.build` (needs nvcc; there is no CPU fallback)."
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Raise NbodyLibError for a non-zero return code."""
    if rc != 0:
        msg = load().nb_error_string(rc).decode("utf-8", "replace")
        raise NbodyLibError(f"libnbody_sm100 error {rc}: {msg}")


def require_cuda():
    """Return torch after checking that a CUDA device and the library are usable."""
    try:
        import torch
    except ImportError as e:  # pragma: no cover - torch is part of the image
        raise ImportError("PyTorch with CUDA is required for the GPU path") from e
    if not torch.cuda.is_available():
        raise ImportError("CUDA device required: no GPU visible (the B200 path has no CPU fallback)")
    load()
    return torch
