"""Force / potential benchmark with the structure of the reference's own
``python -m nbody_streams.fields`` suite (reference fields.py:1199-1600): same
deterministic inputs (``np.random.seed(42)``, standard-normal positions, unit
masses, spline kernel, softening 0.01), the same three GPU cases per precision
(device data without validation, device data with validation, host data with
transfer), the potential, and the net-force check — so its published table
(README.md:781-787: N = 10,240) can be reproduced on a B200.

    python -m nbody_streams_b200.benchmark -N 10240 [--n-warmup 2] [--n-bench 5]

Wall-clock medians with a device synchronise on both sides, as in the
reference.  There is no CPU section: this package has no CPU path.
"""
from __future__ import annotations

import argparse
import time

import numpy as np

from .fields import (_PRECISION_MAP, compute_nbody_forces_gpu, compute_nbody_potential_gpu, get_gpu_info)


def _median_time(fn, n_warmup, n_bench, sync):
    for _ in range(n_warmup):
        out = fn()
        sync()
    times = []
    for _ in range(n_bench):
        sync()
        t0 = time.perf_counter()
        out = fn()
        sync()
        times.append(time.perf_counter() - t0)
    return float(np.median(times)), times, out


def main(argv=None):
    ap = argparse.ArgumentParser(description="N-body GPU benchmark (B200 direct-summation path)")
    ap.add_argument("-N", "--num-particles", type=int, default=25_600)
    ap.add_argument("--n-warmup", type=int, default=2)
    ap.add_argument("--n-bench", type=int, default=5)
    args = ap.parse_args(argv)
    N, n_warmup, n_bench = args.num_particles, args.n_warmup, args.n_bench

    import torch

    info = get_gpu_info()
    print("=" * 80 + "\nN-BODY GPU BENCHMARK (nbody_streams_b200)\n" + "=" * 80)
    if not info["available"]:
        raise SystemExit("No GPU available (this package has no CPU path)")
    print(f"GPU: {info['device_name']}  memory {info['memory_free'] / 1e9:.1f} / {info['memory_total'] / 1e9:.1f} GB free  "
          f"compute capability {info['compute_capability']}")
    print(f"Particles: {N:,}  warm-up {n_warmup}  timed {n_bench}  kernel spline  softening 0.01")

    np.random.seed(42)                                   # the reference's inputs (fields.py:1244-1248)
    pos_f64 = np.random.randn(N, 3)
    mass_f64 = np.ones(N)
    sync = torch.cuda.synchronize
    results = {}
    gi = lambda t: N * N / t / 1e9                       # noqa: E731

    print("\nSECTION 1: GPU FORCE COMPUTATION")
    for precision in ("float32", "float32_kahan", "float64"):
        dtype_np = _PRECISION_MAP[precision][1]
        tdt = torch.float64 if precision == "float64" else torch.float32
        pos_host, mass_host = pos_f64.astype(dtype_np), mass_f64.astype(dtype_np)
        pos_gpu, mass_gpu = torch.from_numpy(pos_host).cuda(), torch.from_numpy(mass_host).cuda()
        h_gpu = torch.full((N,), 0.01, dtype=tdt, device="cuda")
        kw = dict(kernel="spline", precision=precision)
        t_fast, _, acc = _median_time(lambda: compute_nbody_forces_gpu(pos_gpu, mass_gpu, h_gpu, return_cupy=True,
                                                                       skip_validation=True, **kw), n_warmup, n_bench, sync)
        t_val, _, _ = _median_time(lambda: compute_nbody_forces_gpu(pos_gpu, mass_gpu, softening=0.01, return_cupy=True,
                                                                    **kw), n_warmup, n_bench, sync)
        t_xfer, _, _ = _median_time(lambda: compute_nbody_forces_gpu(pos_host, mass_host, softening=0.01, **kw),
                                    n_warmup, n_bench, sync)
        results[precision] = acc.get()
        print(f"[{precision:14s}] device data, no validation: {t_fast * 1e3:8.3f} ms {gi(t_fast):8.1f} Gint/s | with validation: "
              f"{t_val * 1e3:8.3f} ms | host data incl. transfer: {t_xfer * 1e3:8.3f} ms {gi(t_xfer):8.1f} Gint/s")

    print("\nSECTION 2: GPU POTENTIAL COMPUTATION (host arrays in, host array out)")
    for precision in ("float32", "float32_kahan", "float64"):
        dtype_np = _PRECISION_MAP[precision][1]
        pos_host, mass_host = pos_f64.astype(dtype_np), mass_f64.astype(dtype_np)
        t_pot, _, pot = _median_time(lambda: compute_nbody_potential_gpu(pos_host, mass_host, softening=0.01, kernel="spline",
                                                                         precision=precision), n_warmup, n_bench, sync)
        print(f"[{precision:14s}] {t_pot * 1e3:8.3f} ms {gi(t_pot):8.1f} Gint/s   mean potential {pot.mean():.6e}")

    print("\nSECTION 4: NET FORCE (isolated system: sum of m a should vanish to rounding)")
    for precision, acc in results.items():
        eps = np.finfo(np.float64 if precision == "float64" else np.float32).eps
        net = np.linalg.norm(acc.astype(np.float64).sum(0))
        typical = np.median(np.linalg.norm(acc, axis=1))
        rel = net / (typical * N)
        print(f"[{precision:14s}] |sum a| / (N * median|a|) = {rel:.3e}   (machine epsilon {eps:.1e})")
    # accuracy of the float32 arithmetic: float64 evaluation of the SAME float32-rounded inputs as the reference value
    pos_r = pos_f64.astype(np.float32).astype(np.float64)
    ref = compute_nbody_forces_gpu(pos_r, mass_f64, softening=0.01, kernel="spline", precision="float64")
    for precision in ("float32", "float32_kahan"):
        a = results[precision].astype(np.float64)
        print(f"{precision} vs float64 on the same (float32-rounded) inputs: acceleration rel-RMS "
              f"{np.sqrt(((a - ref) ** 2).sum() / (ref ** 2).sum()):.3e}")


if __name__ == "__main__":
    main()
