"""Particle species: the boundary type of ``run_simulation``.

Behavioural mirror of the reference's ``nbody_streams/species.py`` (Species
:24-86, ``_build_particle_arrays`` :93-122, ``_validate_species`` :125-153,
``_split_by_species`` :156-174, ``_emit_performance_warnings`` :177-210): same
field names, validation rules, error types and array conventions, so existing
scripts that build ``Species`` lists run unchanged.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Union

import numpy as np
from numpy.typing import NDArray


class PerformanceWarning(UserWarning):
    """Particle count above the recommended threshold for the chosen backend."""


def _check_per_particle(value, n: int, what: str, owner: str) -> None:
    if np.isscalar(value):
        return
    shape = np.asarray(value).shape
    if shape != (n,):
        raise ValueError(f"Species '{owner}': {what} array shape {shape} != ({n},)")


@dataclass
class Species:
    """One particle species: ``name``, particle count ``N``, ``mass`` and ``softening``
    (each a scalar shared by all N particles or an array of shape (N,))."""

    name: str
    N: int
    mass: Union[float, NDArray]
    softening: Union[float, NDArray] = 0.0

    def __post_init__(self):
        if not isinstance(self.name, str) or not self.name:
            raise ValueError("Species name must be a non-empty string")
        if self.N <= 0:
            raise ValueError(f"Species '{self.name}': N must be > 0, got {self.N}")
        _check_per_particle(self.mass, self.N, "mass", self.name)
        _check_per_particle(self.softening, self.N, "softening", self.name)

    @staticmethod
    def dark(N: int, mass, softening=0.0) -> "Species":
        """Dark-matter species (name ``'dark'``)."""
        return Species(name="dark", N=N, mass=mass, softening=softening)

    @staticmethod
    def star(N: int, mass, softening=0.0) -> "Species":
        """Stellar species (name ``'star'``)."""
        return Species(name="star", N=N, mass=mass, softening=softening)


def _expand(value, n: int) -> NDArray:
    if np.isscalar(value):
        return np.full(n, float(value), dtype=np.float64)
    return np.asarray(value, dtype=np.float64)


def _build_particle_arrays(species: list[Species]) -> tuple[NDArray, NDArray]:
    """Concatenate per-particle float64 mass and softening arrays in species order."""
    masses = [_expand(s.mass, s.N) for s in species]
    soft = [_expand(s.softening, s.N) for s in species]
    return np.concatenate(masses), np.concatenate(soft)


def _validate_species(phase_space: NDArray, species: list[Species]) -> None:
    """ValueError for an empty list, duplicate names or a total N that does not match phase_space."""
    if not species:
        raise ValueError("species list must not be empty")
    names = [s.name for s in species]
    if len(set(names)) != len(names):
        dupes = sorted({n for n in names if names.count(n) > 1})
        raise ValueError(f"Duplicate species names: {dupes}")
    total = sum(s.N for s in species)
    if total != phase_space.shape[0]:
        raise ValueError(
            f"sum(s.N for s in species) = {total} does not match "
            f"phase_space.shape[0] = {phase_space.shape[0]}"
        )


def _split_by_species(xv: NDArray, species: list[Species]) -> dict[str, NDArray]:
    """name -> (N_k, 6) view of the combined phase-space array."""
    out: dict[str, NDArray] = {}
    start = 0
    for s in species:
        out[s.name] = xv[start:start + s.N]
        start += s.N
    return out


def _emit_performance_warnings(N_total: int, architecture: str, method: str) -> None:
    """Thresholds of the reference: >2M any direct, >20k CPU direct, >500k GPU direct."""
    msg = None
    if N_total > 2_000_000 and method != "tree":
        msg = (f"{N_total:,} particles: direct summation at this scale will be extremely slow. "
               "Consider method='tree' for this architecture.")
    elif architecture == "cpu" and method == "direct" and N_total > 20_000:
        msg = (f"{N_total:,} particles with CPU direct summation is O(N^2) and will be very slow. "
               "Consider method='tree' to reduce complexity, or architecture='gpu' "
               "for faster direct summation.")
    elif architecture == "gpu" and method == "direct" and N_total > 500_000:
        msg = (f"{N_total:,} particles with GPU direct summation may be slow at this scale. "
               "Consider method='tree' to improve performance.")
    if msg:
        warnings.warn(msg, PerformanceWarning, stacklevel=4)
