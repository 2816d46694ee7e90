#!/usr/bin/env python3
"""One force(+potential) evaluation for ncu:  profile_one.py N kernel precision mode [reps]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import nbody_streams_b200 as nb  # noqa: E402
from nbody_streams_b200.fields import _PRECISION_MAP, _run_direct  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
kernel = sys.argv[2] if len(sys.argv) > 2 else "plummer"
precision = sys.argv[3] if len(sys.argv) > 3 else "float32_kahan"
mode = {"force": 1, "pot": 2, "both": 3}[sys.argv[4] if len(sys.argv) > 4 else "both"]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
xv, m = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=42069)
prec, dt = _PRECISION_MAP[precision]
pos = torch.from_numpy(xv[:, :3].astype(dt)).cuda()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(reps):
    ev0.record()
    _run_direct(pos, None, float(m[0]), None, 0.01, N, prec, nb.KERNEL_MAP[kernel], mode, nb.G_DEFAULT,
                1e-30 if prec == 2 else 1e-15, dt)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print(f"{kernel} {precision} mode={mode} N={N}: {ms:.3f} ms  {N*N/ms/1e6:.1f} Gint/s")
