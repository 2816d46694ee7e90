#!/usr/bin/env python3
"""Accuracy of the three precisions (and of the reference's own GPU kernels) against the float64 oracle at
full bench size (row subset).  Writes gpurun_out/accuracy.md.  (Step rates: tools/step_rate.py.)"""
import sys
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

Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/accuracy.md").write_text("\n".join(out) + "\n")
print("\n".join(out))
