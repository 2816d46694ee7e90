#!/usr/bin/env python3
"""Build container only (needs /root/reference): is oracle/numba_port.py as fast as the reference's own Numba
functions?  Times the reference's compute_nbody_forces_cpu + compute_nbody_potential_cpu (imported, not copied)
and the restatement on the same inputs with the same number of threads, and compares their results.

    NUMBA_CACHE_DIR=/tmp/nb_numba_cache PYTHONDONTWRITEBYTECODE=1 python tools/numba_vs_reference.py [N]

The table goes into BASELINE.md 2: it is what makes `cpu_baseline_numba` of bench.py (the restatement timed on the
GPU box's host cores) a stand-in for the reference's Numba path there."""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nb_numba_cache")
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

import numpy as np  # noqa: E402


def best(fn, reps=3):
    fn()                                              # JIT / warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


def main():
    from make_golden import import_reference

    import numba

    from nbody_streams_b200.run import make_plummer_sphere
    from oracle import numba_port as P

    fields = import_reference()[0]
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
    xv, m = make_plummer_sphere(N, M_total=1e9, a=1.0, seed=42069)
    G = 4.300917270069976e-06
    print(f"N = {N}, numba {numba.__version__}, {numba.get_num_threads()} threads\n")
    print("| kernel | precision | reference force+potential ms | restatement ms | ratio | max rel diff acc / phi |")
    print("|---|---|---:|---:|---:|---|")
    for prec in ("float32", "float64"):
        dt = np.float32 if prec == "float32" else np.float64
        pos = np.ascontiguousarray(xv[:, :3], dtype=dt)
        mass = np.full(N, m[0], dtype=dt)
        for kernel in ("plummer", "spline"):
            def ref():
                a = fields.compute_nbody_forces_cpu(pos, mass, softening=0.01, G=G, precision=prec, kernel=kernel)
                p = fields.compute_nbody_potential_cpu(pos, mass, softening=0.01, G=G, precision=prec, kernel=kernel)
                return a, p

            def ours():
                a = P.compute_nbody_forces_numba(pos, mass, 0.01, G, kernel, prec)
                p = P.compute_nbody_potential_numba(pos, mass, 0.01, G, kernel, prec)
                return a, p

            t_ref, (a_r, p_r) = best(ref)
            t_our, (a_o, p_o) = best(ours)
            ea = float(np.abs(a_o.astype(np.float64) - a_r).max() / np.abs(a_r).max())
            ep = float(np.abs(p_o.astype(np.float64) - p_r).max() / np.abs(p_r).max())
            print(f"| {kernel} | {prec} | {t_ref * 1e3:.0f} | {t_our * 1e3:.0f} | {t_our / t_ref:.2f} | {ea:.1e} / {ep:.1e} |")


if __name__ == "__main__":
    main()
