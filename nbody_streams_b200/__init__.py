"""nbody_streams_b200 — B200-native direct-summation N-body force engine.

Drop-in for the direct GPU path of appy2806/Nbody_streams: the names below keep
the reference's signatures (``nbody_streams/__init__.py:18-59`` re-exports) and
route to hand-written sm_100a CUDA kernels in ``libnbody_sm100.so`` through a
C ABI (``include/nbody_sm100.h``).  No CPU fallback, no multi-backend dispatch.
"""
from .fields import (G_DEFAULT, KERNEL_MAP, DeviceArray, compute_nbody_forces_gpu,
                     compute_nbody_forces_potential_gpu, compute_nbody_potential_gpu, get_gpu_info)
from .run import make_plummer_sphere, run_nbody_gpu
from .sim import run_simulation
from .species import PerformanceWarning, Species
from .unbind import iterative_unbinding

__version__ = "0.1.0"

__all__ = [
    "Species", "PerformanceWarning", "run_simulation", "run_nbody_gpu", "make_plummer_sphere",
    "compute_nbody_forces_gpu", "compute_nbody_potential_gpu", "compute_nbody_forces_potential_gpu",
    "get_gpu_info", "G_DEFAULT", "KERNEL_MAP", "DeviceArray", "iterative_unbinding",
]
