#!/usr/bin/env python3
"""Second baseline: the reference's own CUDA kernels vs this package on the same B200, same inputs.
Timing protocol of the reference's benchmark (fields.py:1291-1314): 2 warm-ups, median of 5, device
synchronised around each call.  Kernel-only for the reference (its O(N) CuPy glue is not reproduced)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import nbody_streams_b200 as nb  # noqa: E402
import ref_gpu_kernels as ref  # noqa: E402
from nbody_streams_b200.fields import _PRECISION_MAP, _run_direct  # noqa: E402


def med(fn, reps=5, warm=2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for r in range(warm + reps):
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if r >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
xv, m = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=42069)
rows = [f"# Reference GPU kernels vs nbody_streams_b200 on the same B200, Plummer N={N}, scalar softening 0.01\n",
        "Reference = its CUDA-C templates compiled with its own flags (-O3 --use_fast_math) for sm_100, launched as",
        "fields.py does (grid N/128, block 128); kernel time only.  Ours = pack + sweep + finish.\n",
        "| call | precision | kernel | reference ms | reference Gint/s | ours ms | ours Gint/s | speed-up |",
        "|---|---|---|---:|---:|---:|---:|---:|"]
for precision in ("float32", "float32_kahan", "float64"):
    prec, dtn = _PRECISION_MAP[precision]
    dt = torch.float64 if precision == "float64" else torch.float32
    pos_t = torch.from_numpy(xv[:, :3]).to(dt).cuda()
    mass_t = torch.from_numpy(m).to(dt).cuda()
    h_t = torch.full((N,), 0.01, dtype=dt, device="cuda")
    for kernel in ("plummer", "spline"):
        kid = nb.KERNEL_MAP[kernel]
        for kind, mode in (("force", 1), ("potential", 2)):
            t_ref = med(lambda: ref.launch(kind, precision, pos_t, mass_t, h_t, nb.G_DEFAULT, kid),
                        reps=3 if precision == "float64" else 5)
            t_our = med(lambda: _run_direct(pos_t, mass_t, 0.0, h_t, 0.0, N, prec, kid, mode, nb.G_DEFAULT, 1e-15, dtn))
            rows.append(f"| {kind} | {precision} | {kernel} | {t_ref:.3f} | {N*N/t_ref/1e6:.0f} | {t_our:.3f} | "
                        f"{N*N/t_our/1e6:.0f} | {t_ref/t_our:.2f}x |")
            print(rows[-1], flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/ref_gpu_bench.md").write_text("\n".join(rows) + "\n")
