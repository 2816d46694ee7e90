#!/usr/bin/env python3
"""Time the pair-symmetric sweeps against the one-sided ones (pack + sweep + finish, CUDA events).

    python tools/sym_probe.py [--sizes 65536,262144,1000000] [--out gpurun_out/sym_probe.md]

Nothing here is a bench value (see bench.py); it is the A/B that picks nb_sym_min_n and feeds profiles/."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import torch  # noqa: E402

import nbody_streams_b200 as nb  # noqa: E402
from nbody_streams_b200.fields import _PRECISION_MAP, _run_direct  # noqa: E402


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def time_direct(pos_t, m, h, N, prec_key, kernel, mode, algorithm, reps=3):
    prec, dtype_np = _PRECISION_MAP[prec_key]
    kid = nb.KERNEL_MAP[kernel]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms = []
    for r in range(reps + 1):
        ev[0].record()
        _run_direct(pos_t, None, m, None, h, N, prec, kid, mode, nb.G_DEFAULT, 1e-15 if prec != 2 else 1e-30, dtype_np,
                    algorithm=algorithm)
        ev[1].record()
        torch.cuda.synchronize()
        if r:
            ms.append(ev[0].elapsed_time(ev[1]))
    return float(np.median(ms))


def main():
    sizes = [int(x) for x in arg("--sizes", "65536,262144,1000000").split(",")]
    precs = arg("--precs", "float32_kahan,float64").split(",")
    kernels = arg("--kernels", "plummer,spline").split(",")
    out = Path(arg("--out", "gpurun_out/sym_probe.md"))
    lines = ["| N | precision | kernel | mode | one-sided ms | symmetric ms | speed-up | sym Gint/s | x20 flop TF |",
             "|---|---|---|---|---:|---:|---:|---:|---:|"]
    for N in sizes:
        xv, m = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=42069)
        for prec in precs:
            if prec == "float64" and N > 600_000:
                continue
            dt = np.float64 if prec == "float64" else np.float32
            pos_t = torch.from_numpy(np.ascontiguousarray(xv[:, :3].astype(dt))).cuda()
            for kernel in kernels:
                for mode, mname in ((1, "force"), (3, "both"), (2, "pot")):
                    t1 = time_direct(pos_t, float(m[0]), 0.01, N, prec, kernel, mode, "onesided")
                    t2 = time_direct(pos_t, float(m[0]), 0.01, N, prec, kernel, mode, "symmetric")
                    g = N * N / t2 / 1e6
                    row = f"| {N} | {prec} | {kernel} | {mname} | {t1:.3f} | {t2:.3f} | {t1 / t2:.3f} | {g:.0f} | {g * 20 / 1e3:.1f} |"
                    print(row, flush=True)
                    lines.append(row)
    out.parent.mkdir(exist_ok=True)
    out.write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
