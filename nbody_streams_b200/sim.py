"""``run_simulation`` — the reference's top-level entry point, GPU direct leg.

Mirrors reference ``nbody_streams/sim.py:34-471``: same signature, validation,
kwargs handling (``precision`` default ``'float32_kahan'``, unknown kwargs ->
``TypeError``), the hard-coded ``kernel="spline"`` of the GPU-direct branch
(sim.py:404-433) and the ``dict[name -> (N_k, 6)]`` return value.

Only ``architecture='gpu', method='direct'`` is implemented here — that is the
hot path this package replaces.  The CPU and tree legs (Numba / pyfalcon /
libtreeGPU) are out of scope and raise NotImplementedError
pointing at the reference package.
"""
from __future__ import annotations

import warnings
from typing import Literal

import numpy as np
from numpy.typing import NDArray

from .fields import G_DEFAULT
from .run import run_nbody_gpu
from .species import (PerformanceWarning, Species, _build_particle_arrays, _emit_performance_warnings,
                      _split_by_species, _validate_species)

# kwargs the reference pops unconditionally before dispatch (sim.py:324-345)
_DF_KWARGS = ("df_M_sat", "df_coulomb_mode", "df_fixed_ln_lambda", "df_core_gamma", "df_r_core",
              "df_update_interval", "df_shrink_n_iter", "df_shrink_frac", "df_sigma_grid_r",
              "df_apply_radius_factor", "df_sigma_method")


def run_simulation(
    phase_space: np.ndarray,
    species: list[Species],
    time_start: float,
    time_end: float,
    dt: float,
    G: float = G_DEFAULT,
    architecture: Literal["cpu", "gpu"] = "gpu",
    method: Literal["direct", "tree"] = "direct",
    external_potential=None,
    dynamical_friction: bool = False,
    output_dir: str = "./output",
    save_snapshots: bool = True,
    snapshots: int = 100,
    num_files_to_write: int = 1,
    restart_interval: int = 1000,
    continue_run: bool = False,
    overwrite: bool = False,
    verbose: bool = True,
    debug_energy: bool = False,
    **kwargs,
) -> dict[str, NDArray]:
    """Multi-species direct N-body run; returns ``{species name: (N_k, 6) float64}``."""
    if architecture not in ("cpu", "gpu"):
        raise ValueError(f"architecture must be 'cpu' or 'gpu', got '{architecture}'")
    if method not in ("direct", "tree"):
        raise ValueError(f"method must be 'direct' or 'tree', got '{method}'")

    phase_space = np.asarray(phase_space, dtype=np.float64)
    if phase_space.ndim != 2 or phase_space.shape[1] != 6:
        raise ValueError(f"phase_space must be shape (N, 6), got {phase_space.shape}")
    _validate_species(phase_space, species)
    if dynamical_friction and external_potential is None:
        raise ValueError(
            "dynamical_friction=True requires external_potential to be set. "
            "The Chandrasekhar DF formula needs host density and sigma(r) from the external potential."
        )

    mass_arr, softening_arr = _build_particle_arrays(species)
    _emit_performance_warnings(phase_space.shape[0], architecture, method)
    if external_potential is not None and not dynamical_friction:
        m_sat = float(mass_arr.sum())
        if m_sat > 1e10:
            warnings.warn(
                f"Total satellite mass is {m_sat:.2e} M_sun and an external potential is set, but "
                "dynamical_friction=False.  At this mass scale the host-satellite dynamical friction "
                "timescale is short (<~1 Gyr).  Consider setting dynamical_friction=True.",
                PerformanceWarning, stacklevel=2)

    kwargs.pop("theta", None)
    kwargs.pop("nthreads", None)
    external_update_interval = kwargs.pop("external_update_interval", 1)
    precision = kwargs.pop("precision", "float32_kahan")
    force_extra = kwargs.pop("force_extra", None)
    for key in _DF_KWARGS:
        kwargs.pop(key, None)
    # extensions of this package, forwarded to run_nbody_gpu when present
    ext = {k: kwargs.pop(k) for k in ("device", "group", "overlap", "diagnostics", "diagnostics_every", "cuda_graph", "exchange", "async_io", "algorithm")
           if k in kwargs}

    if dynamical_friction:
        raise NotImplementedError(
            "dynamical_friction=True builds its force_extra closure from agama (reference "
            "_chandrasekhar.py), which is outside this package; pass your own force_extra=callable instead.")
    if architecture != "gpu" or method != "direct":
        raise NotImplementedError(
            f"nbody_streams_b200 implements architecture='gpu', method='direct' only "
            f"(got architecture='{architecture}', method='{method}'); use the reference package for the "
            "CPU / tree backends.")
    if kwargs:
        raise TypeError(
            f"run_simulation() got unexpected keyword argument(s) for "
            f"architecture='gpu', method='direct': {sorted(kwargs)}")

    final_xv = run_nbody_gpu(
        phase_space=phase_space, masses=mass_arr, time_start=time_start, time_end=time_end, dt=dt,
        softening=softening_arr, G=G, precision=precision, kernel="spline",
        external_potential=external_potential, external_update_interval=external_update_interval,
        force_extra=force_extra, output_dir=output_dir, save_snapshots=save_snapshots, snapshots=snapshots,
        num_files_to_write=num_files_to_write, restart_interval=restart_interval, continue_run=continue_run,
        overwrite=overwrite, verbose=verbose, debug_energy=debug_energy, species=species, **ext)
    return _split_by_species(final_xv, species)


__all__ = ["run_simulation"]
