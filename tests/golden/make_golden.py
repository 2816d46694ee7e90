#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by running the REFERENCE itself.

Run in the build container only (it needs /root/reference and numba):

    NUMBA_CACHE_DIR=/tmp/nb_numba_cache PYTHONDONTWRITEBYTECODE=1 \
        python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so the outputs of its own CPU
direct-summation path (compute_nbody_forces_cpu / compute_nbody_potential_cpu,
make_plummer_sphere, run_nbody_cpu / run_simulation(architecture='cpu')) on
seeded inputs are committed here as small .npz fixtures.  They pin
  * the C oracle (tests/test_oracle.py),
  * this package's make_plummer_sphere / Species helpers (tests/test_host_logic.py),
  * the CUDA path directly (tests/test_gpu_parity.py).

Nothing is copied from the reference: it is imported and executed.
"""
from __future__ import annotations

import hashlib
import os
import sys
import types
import warnings
from pathlib import Path

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nb_numba_cache")
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True

import numpy as np  # noqa: E402

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
KERNELS = ["newtonian", "plummer", "dehnen_k1", "dehnen_k2", "spline"]


def import_reference():
    """Import nbody_streams.{fields,species,run,sim} without running the package __init__
    (which needs matplotlib / h5py / cupy).  h5py is stubbed: no snapshot is written."""
    warnings.simplefilter("ignore")
    pkg = types.ModuleType("nbody_streams")
    pkg.__path__ = [str(REF / "nbody_streams")]
    sys.modules["nbody_streams"] = pkg
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    import importlib

    fields = importlib.import_module("nbody_streams.fields")
    species = importlib.import_module("nbody_streams.species")
    run = importlib.import_module("nbody_streams.run")
    try:
        sim = importlib.import_module("nbody_streams.sim")
    except Exception as e:  # agama-only pieces may be missing; sim is optional
        print("sim import failed:", e)
        sim = None
    return fields, species, run, sim


def sha16(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    fields, species, run, sim = import_reference()
    G = fields.G_DEFAULT

    # ------------------------------------------------------------------ A
    # all kernels x {float64, float32}, per-particle mass and softening
    rng = np.random.default_rng(1234)
    N = 512
    pos = rng.standard_normal((N, 3))
    mass = rng.lognormal(0.0, 1.0, N)
    h = rng.uniform(0.02, 0.3, N)      # large enough that many pairs are inside h (spline branches)
    out = {"pos": pos, "mass": mass, "h": h, "G": np.float64(G)}
    for k in KERNELS:
        for prec in ("float64", "float32"):
            out[f"acc_{k}_{prec}"] = fields.compute_nbody_forces_cpu(pos, mass, h, G, kernel=k, precision=prec)
            out[f"phi_{k}_{prec}"] = fields.compute_nbody_potential_cpu(pos, mass, h, G, kernel=k, precision=prec)
    np.savez_compressed(OUT / "kernels_n512.npz", **out)

    # ------------------------------------------------------------------ B
    # ragged N, scalar mass + scalar softening, incl. softening = 0 (pure Newtonian limit)
    rng = np.random.default_rng(77)
    N = 301
    pos = rng.standard_normal((N, 3)) * 2.0
    out = {"pos": pos, "G": np.float64(1.0)}
    for k in KERNELS:
        for hs in (0.0, 0.05):
            if hs == 0.0 and k in ("dehnen_k1", "dehnen_k2"):
                pass
            tag = f"{k}_h{hs}"
            out[f"acc_{tag}"] = fields.compute_nbody_forces_cpu(pos, 2.5, hs, 1.0, kernel=k, precision="float64")
            out[f"phi_{tag}"] = fields.compute_nbody_potential_cpu(pos, 2.5, hs, 1.0, kernel=k, precision="float64")
    np.savez_compressed(OUT / "scalar_n301.npz", **out)

    # ------------------------------------------------------------------ C
    # the reference's own known-answer inputs:
    #   tests/test_forces_versus_cpu.py:6-10  (N=100, seed 42, spline, h=0.01, float64)
    #   tests/test_newtons_third_law.py:24-31 (N=1000, seed 42, equal masses, h=0.01, spline)
    np.random.seed(42)
    pos100 = np.random.randn(100, 3).astype(np.float64)
    acc100 = fields.compute_nbody_forces_cpu(pos100, np.ones(100), softening=0.01, kernel="spline",
                                             precision="float64")
    np.random.seed(42)
    pos1000 = np.random.randn(1000, 3)
    acc1000 = fields.compute_nbody_forces_cpu(pos1000, np.ones(1000), softening=0.01, kernel="spline",
                                              precision="float64")
    np.savez_compressed(OUT / "reference_tests.npz", pos100=pos100, acc100=acc100, pos1000=pos1000,
                        acc1000=acc1000)

    # ------------------------------------------------------------------ D
    # make_plummer_sphere pins (reference run.py:1225-1325)
    out = {}
    for tag, kw in {
        "n1000_default": dict(N=1000),
        "n2000_c2star": dict(N=2000, M_total=1e7, a=0.3, seed=42070),
        "n10000_c1": dict(N=10_000, M_total=1e9, a=1.0, seed=42069),
    }.items():
        xv, m = run.make_plummer_sphere(**kw)
        if kw["N"] <= 2000:
            out[f"xv_{tag}"] = xv
        else:
            out[f"xv_head_{tag}"] = xv[:64]
            out[f"xv_tail_{tag}"] = xv[-64:]
        out[f"sha_{tag}"] = np.array(sha16(xv))
        out[f"m0_{tag}"] = np.float64(m[0])
    np.savez_compressed(OUT / "plummer_ic.npz", **out)

    # ------------------------------------------------------------------ E
    # BASELINE config 1: Plummer N=10k, spline h=0.05, float64, 100 KDK steps dt=1e-4 on the CPU path
    xv, m = run.make_plummer_sphere(10_000, M_total=1e9, a=1.0, seed=42069)
    acc0 = fields.compute_nbody_forces_cpu(xv[:, :3], m, 0.05, G, kernel="spline", precision="float64")
    phi0 = fields.compute_nbody_potential_cpu(xv[:, :3], m, 0.05, G, kernel="spline", precision="float64")
    acc0_pl = fields.compute_nbody_forces_cpu(xv[:, :3], m, 0.05, G, kernel="plummer", precision="float64")
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        sp = [species.Species.dark(10_000, mass=1e5, softening=0.05)]
        if sim is not None:
            res = sim.run_simulation(xv, sp, 0.0, 100 * 1e-4, 1e-4, architecture="cpu", method="direct",
                                     output_dir=td, save_snapshots=False, verbose=False)
            final = res["dark"]
        else:
            final = run.run_nbody_cpu(xv, m, 0.0, 100 * 1e-4, 1e-4, 0.05, G=G, method="direct", kernel="spline",
                                      output_dir=td, save_snapshots=False, verbose=False)
    phi1 = fields.compute_nbody_potential_cpu(final[:, :3], m, 0.05, G, kernel="spline", precision="float64")
    ke0 = 0.5 * np.dot(m, np.sum(xv[:, 3:] ** 2, axis=1))
    ke1 = 0.5 * np.dot(m, np.sum(final[:, 3:] ** 2, axis=1))
    np.savez_compressed(
        OUT / "c1_plummer10k.npz",
        acc0=acc0, phi0=phi0, acc0_plummer_rows=acc0_pl[::10],
        final_rows=final[::10], ke0=ke0, pe0=0.5 * np.dot(m, phi0), ke1=ke1, pe1=0.5 * np.dot(m, phi1),
        p0=m @ xv[:, 3:], p1=m @ final[:, 3:],
    )

    # ------------------------------------------------------------------ F
    # multi-species (dark + star with per-particle softening + 1 heavy BH), float32 and float64
    xd, md = run.make_plummer_sphere(1000, M_total=1e9, a=1.0, seed=1)
    xs, ms = run.make_plummer_sphere(250, M_total=2e7, a=0.3, seed=2)
    xb = np.zeros((1, 6))
    xvm = np.vstack([xd, xs, xb])
    sp = [species.Species.dark(1000, mass=float(md[0]), softening=0.05),
          species.Species.star(250, mass=float(ms[0]), softening=np.linspace(0.01, 0.03, 250)),
          species.Species("bh", 1, 1e8, softening=0.001)]
    mass_arr, soft_arr = species._build_particle_arrays(sp)
    out = {"xv": xvm, "mass": mass_arr, "h": soft_arr}
    for prec in ("float64", "float32"):
        out[f"acc_spline_{prec}"] = fields.compute_nbody_forces_cpu(xvm[:, :3], mass_arr, soft_arr, G,
                                                                  kernel="spline", precision=prec)
        out[f"phi_spline_{prec}"] = fields.compute_nbody_potential_cpu(xvm[:, :3], mass_arr, soft_arr, G,
                                                                     kernel="spline", precision=prec)
        out[f"acc_plummer_{prec}"] = fields.compute_nbody_forces_cpu(xvm[:, :3], mass_arr, soft_arr, G,
                                                                   kernel="plummer", precision=prec)
    with tempfile.TemporaryDirectory() as td:
        if sim is not None:
            res = sim.run_simulation(xvm, sp, 0.0, 20 * 1e-4, 1e-4, architecture="cpu", method="direct",
                                     output_dir=td, save_snapshots=False, verbose=False)
            out["final_dark"] = res["dark"]
            out["final_star"] = res["star"]
            out["final_bh"] = res["bh"]
    np.savez_compressed(OUT / "multispecies_n1251.npz", **out)

    for f in sorted(OUT.glob("*.npz")):
        print(f"{f.name:32s} {f.stat().st_size/1024:8.1f} KiB")


if __name__ == "__main__":
    main()
