#!/usr/bin/env python3
"""Time one force evaluation over a list of N with both sweeps (one-sided / pair-symmetric): where is the
crossover that nb_sym_min_n() should encode, and are there size-dependent anomalies?

    python tools/n_sweep.py [kernel] [h] [N,N,...] [precision]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import nbody_streams_b200 as nb  # noqa: E402
from nbody_streams_b200.fields import _PRECISION_MAP, _run_direct  # noqa: E402

kernel = sys.argv[1] if len(sys.argv) > 1 else "spline"
h = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
Ns = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4096, 8192, 10000, 16384, 24576, 32768, 40000,
                                                                        49152, 65536, 100000, 131072, 200000, 262144]
precision = sys.argv[4] if len(sys.argv) > 4 else "float32_kahan"
prec, dt = _PRECISION_MAP[precision]
xv, m = nb.make_plummer_sphere(max(Ns), M_total=1e9, a=1.0, seed=1)
print(f"| N | {precision} {kernel} force: one-sided ms | symmetric ms | one-sided Gint/s | symmetric Gint/s | sym/one |")
print("|---:|---:|---:|---:|---:|---:|")
for N in Ns:
    pos = torch.from_numpy(np.ascontiguousarray(xv[:N, :3].astype(dt))).cuda()
    res = {}
    for alg in ("onesided", "symmetric"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for r in range(6):
            e0.record()
            _run_direct(pos, None, float(m[0]), None, h, N, prec, nb.KERNEL_MAP[kernel], 1, nb.G_DEFAULT,
                        1e-30 if prec == 2 else 1e-15, dt, algorithm=alg)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[alg] = float(np.median(ts[2:]))
    a, b = res["onesided"], res["symmetric"]
    print(f"| {N} | {a:.4f} | {b:.4f} | {N * N / a / 1e6:.0f} | {N * N / b / 1e6:.0f} | {a / b:.3f} |", flush=True)
