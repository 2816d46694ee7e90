#!/usr/bin/env python3
"""Time one force evaluation for a list of N (looking for size-dependent anomalies)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import nbody_streams_b200 as nb
from nbody_streams_b200.fields import _PRECISION_MAP, _run_direct

kernel = sys.argv[1] if len(sys.argv) > 1 else "spline"
h = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
Ns = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [40000, 65536, 98304, 100000, 100352, 131072, 200000, 262144]
prec, dt = _PRECISION_MAP["float32_kahan"]
xv, m = nb.make_plummer_sphere(max(Ns), M_total=1e9, a=1.0, seed=1)
for N in Ns:
    pos = torch.from_numpy(xv[:N, :3].astype(dt)).cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for r in range(4):
        e0.record()
        _run_direct(pos, None, float(m[0]), None, h, N, prec, nb.KERNEL_MAP[kernel], 1, nb.G_DEFAULT, 1e-15, dt)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[1:]))
    print(f"{kernel} h={h} N={N:7d}: {ms:9.3f} ms  {N*N/ms/1e6:8.1f} Gint/s", flush=True)
