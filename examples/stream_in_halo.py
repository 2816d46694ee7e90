#!/usr/bin/env python3
"""A dwarf-galaxy progenitor dissolving into a stream inside a static host halo — the use case the
reference package is named after (its examples/stream_Nbody.ipynb), on the B200 path:

  * self-gravity: direct summation on the GPU (run_simulation -> run_nbody_gpu, spline softening)
  * host halo: an analytic logarithmic potential evaluated ON THE DEVICE (an object with
    ``accepts_device_arrays = True`` and ``force(pos, t=)``; the reference's PotentialGPU objects play
    this role), so the state never leaves the GPU during the run
  * afterwards: which particles are still bound (iterative_unbinding, direct_gpu)

    python examples/stream_in_halo.py [N] [steps]
"""
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nbody_streams_b200 as nb  # noqa: E402


class LogHaloOnDevice:
    """Phi = v0^2/2 ln(r^2 + rc^2); acceleration -v0^2 x / (r^2 + rc^2).  Works on device arrays."""

    accepts_device_arrays = True

    def __init__(self, v0=220.0, rc=5.0):
        self.v0sq, self.rc2 = v0 * v0, rc * rc

    def force(self, pos, t=0.0):
        import torch

        p = pos.as_subclass(torch.Tensor)
        return -self.v0sq * p / ((p * p).sum(1, keepdim=True) + self.rc2)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000
    # progenitor: 1e7 Msun Plummer sphere, scale radius 0.2 kpc, on an eccentric orbit starting at 25 kpc
    xv, m = nb.make_plummer_sphere(n, M_total=1e7, a=0.2, seed=7)
    xv[:, :3] += [25.0, 0.0, 0.0]
    xv[:, 3:] += [0.0, 90.0, 20.0]
    dt = 2.5e-4                                       # kpc/(km/s) = 0.978 Gyr: dt ~ 0.24 Myr
    dark = nb.Species.dark(n, mass=float(m[0]), softening=0.02)
    with tempfile.TemporaryDirectory() as out:
        t0 = time.perf_counter()
        res = nb.run_simulation(xv, [dark], 0.0, steps * dt, dt, architecture="gpu", method="direct",
                                external_potential=LogHaloOnDevice(), output_dir=out, save_snapshots=False,
                                verbose=False)
        wall = time.perf_counter() - t0
    final = res["dark"]
    print(f"N = {n}, {steps} KDK steps ({steps * dt * 0.978:.2f} Gyr) in {wall:.2f} s wall "
          f"(includes ~1 s of one-off set-up and the final restart file; steady state is ~0.2 ms/step at N = 20k)")
    (bound,), cpos, cvel = nb.iterative_unbinding(final[:, :3], final[:, 3:], m, potential_compute_method="direct_gpu",
                                                  softening=0.02, verbose=False)
    r = np.linalg.norm(final[:, :3] - cpos, axis=1)
    print(f"remnant at {np.around(cpos, 2)} kpc, bound fraction {bound.mean():.3f}, "
          f"stream extent (99th percentile distance from the remnant) {np.percentile(r, 99):.1f} kpc")


if __name__ == "__main__":
    main()
