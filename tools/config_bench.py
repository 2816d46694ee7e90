#!/usr/bin/env python3
"""Throughput of one force(+potential) evaluation for the BASELINE configs that are not the bench line
(C2 two-species 120k, C4 Plummer 262k float64, C5 2M+1 three-species) on one GPU."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import nbody_streams_b200 as nb
from nbody_streams_b200.fields import _PRECISION_MAP, _run_direct


def timeit(pos, mass, h, precision, kernel, mode, reps=3):
    prec, dt = _PRECISION_MAP[precision]
    N = pos.shape[0]
    pos_t = torch.from_numpy(np.ascontiguousarray(pos.astype(dt))).cuda()
    m_t = None if np.isscalar(mass) else torch.from_numpy(mass.astype(dt)).cuda()
    h_t = None if np.isscalar(h) else torch.from_numpy(h.astype(dt)).cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for r in range(reps + 1):
        e0.record()
        _run_direct(pos_t, m_t, mass if m_t is None else 0.0, h_t, h if h_t is None else 0.0, N, prec,
                    nb.KERNEL_MAP[kernel], mode, nb.G_DEFAULT, 1e-30 if prec == 2 else 1e-15, dt)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[1:]))
    return ms, N * N / ms / 1e6


rows = ["| config | N | precision | kernel | mode | ms | Gint/s |", "|---|---:|---|---|---|---:|---:|"]
xd, md = nb.make_plummer_sphere(100_000, M_total=1e9, a=1.0, seed=42069)
xs, ms_ = nb.make_plummer_sphere(20_000, M_total=1e7, a=0.3, seed=42070)
pos = np.vstack([xd[:, :3], xs[:, :3]]); mass = np.concatenate([np.full(100_000, 1e4), np.full(20_000, 500.0)])
h = np.concatenate([np.full(100_000, 0.05), np.linspace(0.01, 0.03, 20_000)])
for kernel in ("spline", "plummer"):
    for mode, mn in ((1, "force"), (3, "force+pot")):
        ms, g = timeit(pos, mass, h, "float32", kernel, mode)
        rows.append(f"| C2 dark 100k + star 20k (per-particle h) | {len(pos)} | float32 | {kernel} | {mn} | {ms:.3f} | {g:.0f} |")
xv, m = nb.make_plummer_sphere(262_144, M_total=1e9, a=1.0, seed=42069)
for kernel in ("spline", "plummer"):
    for mode, mn in ((1, "force"), (3, "force+pot")):
        ms, g = timeit(xv[:, :3], float(m[0]), 0.02, "float64", kernel, mode, reps=2)
        rows.append(f"| C4 Plummer 262k | 262144 | float64 | {kernel} | {mn} | {ms:.3f} | {g:.0f} |")
xd, md = nb.make_plummer_sphere(1_600_000, M_total=1e9, a=1.0, seed=1)
xs, ms_ = nb.make_plummer_sphere(400_000, M_total=2e7, a=0.3, seed=2)
pos = np.vstack([xd[:, :3], xs[:, :3], np.zeros((1, 3))])
mass = np.concatenate([md, ms_, [1e8]]); h = np.concatenate([np.full(1_600_000, 0.05), np.full(400_000, 0.02), [0.001]])
for mode, mn in ((1, "force"), (3, "force+pot")):
    ms, g = timeit(pos, mass, h, "float32", "spline", mode, reps=2)
    rows.append(f"| C5 dark 1.6M + star 400k + BH | {len(pos)} | float32 | spline | {mn} | {ms:.3f} | {g:.0f} |")
print("\n".join(rows))
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/config_bench.md").write_text("# One-GPU throughput of the other BASELINE configurations (round 2: pair-symmetric sweep; round 1 numbers in r01_config_bench_c2_c4_c5.md)\n\n" + "\n".join(rows) + "\n")
