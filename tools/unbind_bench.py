#!/usr/bin/env python3
"""Time iterative unbinding on a stripped satellite (GPU box): the device-resident N x N_bound loop of
nbody_streams_b200.unbind against the reference's loop structure (utils/main.py:2034-2037: a full N x N
potential call with masked masses and host arrays every iteration) driven through the same potential API.

    python tools/unbind_bench.py [N] > gpurun_out/unbind_bench.txt
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nbody_streams_b200 as nb  # noqa: E402
from nbody_streams_b200.unbind import iterative_unbinding  # noqa: E402


def satellite(n, seed=5):
    rng = np.random.default_rng(seed)
    xv, m = nb.make_plummer_sphere(n, M_total=1e9, a=1.0, seed=seed)
    pick = rng.random(n) < 0.3
    xv[pick, :3] *= rng.uniform(1.0, 8.0, pick.sum())[:, None]
    xv[pick, 3:] *= rng.uniform(0.8, 4.0, pick.sum())[:, None]
    return xv, m


def reference_loop(xv, m, softening, cpos, cvel, iters=50, tol=1e-4):
    """The reference's loop, verbatim in structure, calling compute_nbody_potential_gpu with host arrays."""
    pos_rel, vel_rel = xv[:, :3] - cpos, xv[:, 3:] - cvel
    mask = np.ones(len(xv), dtype=bool)
    n_it = 0
    for _ in range(iters):
        phi = nb.compute_nbody_potential_gpu(pos_rel, m * mask, softening=softening)
        bound = (phi + 0.5 * np.sum(vel_rel ** 2, axis=1)) < 0
        frac = np.mean(bound != mask)
        mask = bound
        n_it += 1
        if frac < tol:
            break
    return mask, n_it


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    xv, m = satellite(n)
    kw = dict(potential_compute_method="direct_gpu", softening=0.02, verbose=False)
    # warm-up (library load, scratch allocation)
    iterative_unbinding(xv[:4096, :3], xv[:4096, 3:], m[:4096], **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    (b, hist), cp, cv = iterative_unbinding(xv[:, :3], xv[:, 3:], m, return_history=True, **kw)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    n_it = len(hist)
    print(f"N = {n}: bound {int(b.sum())} ({b.mean():.3f}), {n_it} iterations + 1 centring potential")
    ours = t1 - t0
    print(f"  this library (device-resident, N x N_bound sweeps): {ours:8.3f} s")
    t0 = time.perf_counter()
    phi0 = nb.compute_nbody_potential_gpu(xv[:, :3], m, softening=0.02)          # centring call of the reference
    mask, n_ref = reference_loop(xv, m, 0.02, cp, cv)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"  reference loop structure on the same kernels (N x N per iteration, host arrays): {t2 - t0:8.3f} s "
          f"({n_ref} iterations); masks equal: {np.array_equal(mask, b.astype(bool))}")
    print(f"  ratio {(t2 - t0) / ours:.2f}x")


if __name__ == "__main__":
    main()
