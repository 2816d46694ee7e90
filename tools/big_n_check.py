#!/usr/bin/env python3
"""Large-N sanity of the pair-symmetric sweep (index arithmetic beyond 2^31 pairs per block row, 128 MB workspace):
N = 4M Plummer, float32_kahan spline force+potential, oracle rows at the end of the range, Newton's third law.

    python tools/big_n_check.py"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import nbody_streams_b200 as nb, oracle
oracle.build()
N = 4_000_000
xv, m = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=3)
pos = xv[:, :3]
t0 = time.perf_counter()
acc, phi = nb.compute_nbody_forces_potential_gpu(pos, float(m[0]), 0.01, precision="float32_kahan", kernel="spline")
print("N=4M spline both e2e s:", time.perf_counter() - t0)
rows = (3_999_000, 4_000_000)
p64 = pos.astype(np.float32).astype(np.float64)
ref = oracle.compute_nbody_forces_cpu(p64, float(np.float32(m[0])), float(np.float32(0.01)), kernel="spline", rows=rows, fast=True)
pref = oracle.compute_nbody_potential_cpu(p64, float(np.float32(m[0])), float(np.float32(0.01)), kernel="spline", rows=rows, fast=True)
a = acc[rows[0]:rows[1]].astype(np.float64); p = phi[rows[0]:rows[1]].astype(np.float64)
print("acc rel-RMS", np.sqrt(((a-ref)**2).sum()/(ref**2).sum()), "phi rel-RMS", np.sqrt(((p-pref)**2).sum()/(pref**2).sum()), "finite", np.isfinite(acc).all())
f = acc.astype(np.float64) * m[:, None]
print("newton3", np.linalg.norm(f.sum(0)) / (np.median(np.linalg.norm(f, axis=1)) * N))
