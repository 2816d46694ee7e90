#!/usr/bin/env python3
"""Golden vectors for iterative unbinding, produced by running the REFERENCE's own
``utils.iterative_unbinding(potential_compute_method='direct')`` (reference utils/main.py:1722-2079;
its Numba CPU potential, float64, spline) on seeded satellites with a stripped, unbound tail.

Build container only (needs /root/reference and numba):

    NUMBA_CACHE_DIR=/tmp/nb_numba_cache PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_unbinding.py

Writes tests/golden/unbinding.npz.  Nothing is copied from the reference: it is imported and run.
"""
from __future__ import annotations

import importlib
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


def import_reference():
    warnings.simplefilter("ignore")
    for name, path in (("nbody_streams", REF / "nbody_streams"), ("nbody_streams.utils", REF / "nbody_streams" / "utils")):
        pkg = types.ModuleType(name)
        pkg.__path__ = [str(path)]
        sys.modules[name] = pkg
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    run = importlib.import_module("nbody_streams.run")
    umain = importlib.import_module("nbody_streams.utils.main")
    return run, umain


def satellite(run, n_dark, n_star, seed):
    """Plummer dark halo + compact stellar component on a bulk orbit, a quarter of the dark and a tenth
    of the stars pushed outwards and sped up (an unbound tidal tail)."""
    rng = np.random.default_rng(seed)
    xd, md = run.make_plummer_sphere(n_dark, M_total=1e8, a=0.5, seed=seed + 1)
    xs, ms = run.make_plummer_sphere(n_star, M_total=2e6, a=0.15, seed=seed + 2)
    for xv, frac in ((xd, 0.25), (xs, 0.10)):
        pick = rng.random(len(xv)) < frac
        xv[pick, :3] *= rng.uniform(1.0, 6.0, pick.sum())[:, None]
        xv[pick, 3:] *= rng.uniform(0.8, 4.0, pick.sum())[:, None]
        xv[pick, 3:] += rng.normal(0.0, 5.0, (pick.sum(), 3))
    bulk = np.array([20.0, -5.0, 3.0, 100.0, 50.0, -30.0])
    return xd + bulk, md, xs + bulk, ms


def main():
    run, umain = import_reference()
    out = {}

    xd, md, xs, ms = satellite(run, 3000, 600, 100)
    out.update(a_xd=xd, a_md=md, a_xs=xs, a_ms=ms)
    # A: two components, automatic centre on the dark particles, history returned
    (bd, bs, hd, hs), cp, cv = umain.iterative_unbinding(
        xd[:, :3], xd[:, 3:], md, xs[:, :3], xs[:, 3:], ms, potential_compute_method="direct", softening=0.03,
        center_on="dark", vel_aperture=5.0, verbose=False, return_history=True)
    out.update(a_bound_dark=bd, a_bound_star=bs, a_hist_dark=np.array(hd), a_hist_star=np.array(hs), a_cpos=cp, a_cvel=cv)
    # B: same particles, centre on both components, small aperture, larger top_fraction, scalar masses
    (bd, bs), cp, cv = umain.iterative_unbinding(
        xd[:, :3], xd[:, 3:], float(md[0]), xs[:, :3], xs[:, 3:], float(ms[0]), potential_compute_method="direct",
        softening=0.05, center_on="both", vel_aperture=0.3, verbose=False, top_fraction=0.05)
    out.update(b_bound_dark=bd, b_bound_star=bs, b_cpos=cp, b_cvel=cv)
    # C: dark only, centre given, at most 3 iterations (not converged), ragged N
    xd2, md2, _, _ = satellite(run, 1237, 10, 200)
    out.update(c_xd=xd2, c_md=md2)
    (bd,), cp, cv = umain.iterative_unbinding(
        xd2[:, :3], xd2[:, 3:], md2, center_position=[20.0, -5.0, 3.0], center_velocity=[100.0, 50.0, -30.0],
        recursive_iter_converg=3, potential_compute_method="direct", softening=0.03, verbose=False)
    out.update(c_bound_dark=bd, c_cpos=cp, c_cvel=cv)

    np.savez_compressed(OUT / "unbinding.npz", **out)
    f = OUT / "unbinding.npz"
    print(f"{f.name} {f.stat().st_size/1024:.1f} KiB;  A: bound dark {out['a_bound_dark'].sum()}/{3000}, star "
          f"{out['a_bound_star'].sum()}/600 in {len(out['a_hist_dark'])} iterations;  B: {out['b_bound_dark'].sum()}, "
          f"{out['b_bound_star'].sum()};  C: {out['c_bound_dark'].sum()}/1237")


if __name__ == "__main__":
    main()
