#!/usr/bin/env python3
"""Per-step time of run_nbody_gpu over N (GPU box): `python tools/step_rate.py > gpurun_out/step_rate.md`.

Per-step time = (t(3S) - t(S)) / 2S with S chosen so that S steps take >= ~1 s, each t the minimum of 3 runs:
the difference removes set-up, the initial force evaluation and the final restart.npz write that every run pays
once.  Wall clock around the public call (host arrays in, host arrays out)."""
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nbody_streams_b200 as nb  # noqa: E402


def timed(xs, ms, steps, kw):
    best = float("inf")
    for _ in range(3):
        with tempfile.TemporaryDirectory() as td:
            t0 = time.perf_counter()
            nb.run_nbody_gpu(xs, ms, 0.0, steps * 1e-4, 1e-4, 0.05, output_dir=td, **kw)
            best = min(best, time.perf_counter() - t0)
    return best


def main():
    print("| N | precision | kernel | CUDA graph | S | ms/step | Gint/s |")
    print("|---:|---|---|---|---:|---:|---:|")
    Ns = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else (1_000, 10_000, 40_000, 120_000, 262_144)
    for N in Ns:
        xs, ms = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=1)
        for prec, kernel, graph in (("float32_kahan", "spline", None), ("float32_kahan", "spline", False),
                                    ("float64", "spline", None)):
            if N > 131_072 and graph is False:
                continue                      # by default graphs are used up to N = 131072 only: same path
            kw = dict(save_snapshots=False, verbose=False, precision=prec, kernel=kernel, cuda_graph=graph)
            pilot = max(timed(xs, ms, 60, kw) - timed(xs, ms, 20, kw), 1e-6) / 40
            S = int(min(max(1.0 / pilot, 20), 30_000))
            per = (timed(xs, ms, 3 * S, kw) - timed(xs, ms, S, kw)) / (2 * S)
            print(f"| {N} | {prec} | {kernel} | {'yes' if graph is None and N <= 131_072 else 'no'} | {S} | {per * 1e3:.4f} | "
                  f"{N * N / per / 1e9:.1f} |", flush=True)


if __name__ == "__main__":
    main()
