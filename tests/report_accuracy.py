#!/usr/bin/env python3
"""Accuracy of the three precisions against the float64 oracle at full bench size (row subset), and
small-N step rate of run_nbody_gpu.  Writes gpurun_out/accuracy.md."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import nbody_streams_b200 as nb  # noqa: E402
import oracle  # noqa: E402  (lives under tests/: the oracle is only the checker here)
import ref_gpu_kernels as refk  # noqa: E402


def rel_rms(a, r):
    a, r = np.asarray(a, np.float64), np.asarray(r, np.float64)
    return float(np.sqrt(((a - r) ** 2).sum() / (r ** 2).sum()))


def rel_max(a, r):
    a, r = np.asarray(a, np.float64), np.asarray(r, np.float64)
    return float((np.linalg.norm(a - r, axis=1) / np.linalg.norm(r, axis=1)).max())


out = ["# Accuracy vs the float64 oracle, Plummer N=1M (a=1, M=1e9, softening 0.01), rows [500000, 502048)\n",
       "Reference values: oracle float64 on the SAME float64 inputs (so float32 numbers include input rounding).",
       "`cpu f32` = the oracle's restatement of the reference CPU float32 path (float pair math, double sums).\n",
       "| kernel | precision | acc rel-RMS | acc max rel | phi rel-RMS |", "|---|---|---:|---:|---:|"]
xv, m = nb.make_plummer_sphere(1_000_000, M_total=1e9, a=1.0, seed=42069)
pos = xv[:, :3]
rows = (500_000, 502_048)
sl = slice(*rows)
for kernel in ("plummer", "spline"):
    ref = oracle.compute_nbody_forces_cpu(pos, m, 0.01, kernel=kernel, rows=rows, fast=True)
    pref = oracle.compute_nbody_potential_cpu(pos, m, 0.01, kernel=kernel, rows=rows, fast=True)
    c32 = oracle.compute_nbody_forces_cpu(pos, m, 0.01, kernel=kernel, rows=rows, fast=True, precision="float32")
    p32 = oracle.compute_nbody_potential_cpu(pos, m, 0.01, kernel=kernel, rows=rows, fast=True, precision="float32")
    out.append(f"| {kernel} | cpu f32 (reference CPU path) | {rel_rms(c32, ref):.2e} | {rel_max(c32, ref):.2e} | {rel_rms(p32, pref):.2e} |")
    for prec in ("float32", "float32_kahan", "float64"):
        acc, phi = nb.compute_nbody_forces_potential_gpu(pos, m, 0.01, precision=prec, kernel=kernel)
        out.append(f"| {kernel} | gpu {prec} | {rel_rms(acc[sl], ref):.2e} | {rel_max(acc[sl], ref):.2e} | {rel_rms(phi[sl], pref):.2e} |")
        if refk.available():   # the reference's own GPU kernel of the same precision, same inputs
            import torch

            dt = torch.float64 if prec == "float64" else torch.float32
            pos_t = torch.from_numpy(pos).to(dt).cuda()
            m_t = torch.from_numpy(m).to(dt).cuda()
            h_t = torch.full((len(m),), 0.01, dtype=dt, device="cuda")
            ra = refk.launch("force", prec, pos_t, m_t, h_t, nb.G_DEFAULT, nb.KERNEL_MAP[kernel]).cpu().numpy()
            rp = refk.launch("potential", prec, pos_t, m_t, h_t, nb.G_DEFAULT, nb.KERNEL_MAP[kernel]).cpu().numpy()
            out.append(f"| {kernel} | reference GPU kernel {prec} | {rel_rms(ra[sl], ref):.2e} | {rel_max(ra[sl], ref):.2e} | {rel_rms(rp[sl], pref):.2e} |")

out.append("\n# Step rate of run_nbody_gpu (spline, float32_kahan, no snapshots)\n")
out.append("Per-step time = (t(3S steps) - t(S steps)) / 2S, which removes set-up, the initial force evaluation and the")
out.append("final restart.npz write (np.savez_compressed, ~3 ms per 10k particles) that every run pays once.\n")
out.append("| N | S | ms/step | Gint/s | one-off overhead ms |")
out.append("|---:|---:|---:|---:|---:|")
for N, steps in ((1_000, 300), (10_000, 300), (40_000, 100), (100_000, 40), (262_144, 10)):
    xs, ms_ = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=1)
    kw = dict(save_snapshots=False, verbose=False, output_dir="/tmp/acc_out")
    nb.run_nbody_gpu(xs, ms_, 0.0, 5e-4, 1e-4, 0.05, **kw)
    ts = []
    for k in (1, 3):
        t0 = time.perf_counter()
        nb.run_nbody_gpu(xs, ms_, 0.0, k * steps * 1e-4, 1e-4, 0.05, **kw)
        ts.append(time.perf_counter() - t0)
    per = (ts[1] - ts[0]) / (2 * steps)
    out.append(f"| {N} | {steps} | {per * 1e3:.3f} | {N * N / per / 1e9:.1f} | {(ts[0] - steps * per) * 1e3:.1f} |")
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/accuracy.md").write_text("\n".join(out) + "\n")
print("\n".join(out))
