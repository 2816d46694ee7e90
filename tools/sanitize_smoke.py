#!/usr/bin/env python3
"""Small pass over every kernel of libnbody_sm100.so (GPU box): `python tools/sanitize_smoke.py`.
Written for compute-sanitizer memcheck, which this GPU pool does not allow; out-of-bounds writes are
checked instead by the guard-band test tests/test_gpu_guard.py.

All 5 softening kernels x 3 precisions x {force, potential, both} at ragged N with per-particle mass and
softening (generic tiles), uniform ones (fast-path tiles), n_sources, a short KDK run (kick/drift/pack/energy)
and an unbinding call.  Sizes are tiny: the point is addressing, not speed.
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nbody_streams_b200 as nb  # noqa: E402
from nbody_streams_b200 import fields  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    for n in (1, 255, 777, 1500):
        pos = rng.standard_normal((n, 3))
        mass = rng.uniform(0.5, 2.0, n)
        h = rng.uniform(0.01, 0.2, n)
        for prec in ("float32", "float32_kahan", "float64"):
            for kernel in fields.KERNEL_MAP:
                nb.compute_nbody_forces_gpu(pos, mass, h, precision=prec, kernel=kernel)
                nb.compute_nbody_potential_gpu(pos, mass, h, precision=prec, kernel=kernel)
                a, p = nb.compute_nbody_forces_potential_gpu(pos, 1.0, 0.05, precision=prec, kernel=kernel)
                assert np.isfinite(a).all() and np.isfinite(p).all()
            nb.compute_nbody_potential_gpu(pos, mass, h, precision=prec, n_sources=n // 2)
    xv, m = nb.make_plummer_sphere(1500, M_total=1e9, a=1.0, seed=1)
    for prec in ("float32_kahan", "float64"):
        with tempfile.TemporaryDirectory() as td:
            out = nb.run_nbody_gpu(xv, m, 0.0, 5e-4, 1e-4, 0.05, precision=prec, kernel="spline", output_dir=td,
                                   save_snapshots=False, verbose=False, debug_energy=True)
            assert np.isfinite(out).all()
    (b,), _, _ = nb.iterative_unbinding(xv[:, :3], 3.0 * xv[:, 3:], m, potential_compute_method="direct_gpu",
                                        softening=0.05, verbose=False)
    print("sanitize smoke done; bound", int(b.sum()))


if __name__ == "__main__":
    main()
