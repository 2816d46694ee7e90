"""numba_port.py — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

A Numba restatement of the reference's CPU direct-summation path in the reference's own technology
(`@njit(parallel=True, fastmath=True)`, `prange` over the i-particles, `if i == j: continue`, the max-softening
rule, float64 running sums), so that `bench.py` can time "the Numba path on the box's own host cores" — the
reference itself is Python and does not travel to the GPU box.  Restates (never copies)

    _get_force_kernel       nbody_streams/fields.py:122-189
    _compute_forces_cpu     nbody_streams/fields.py:191-256   (decorators :191: parallel, fastmath, cache)
    _get_potential_kernel   nbody_streams/fields.py:258-314
    _compute_potential_cpu  nbody_streams/fields.py:317-369

from oracle/nbody_oracle.c (same arithmetic, same typing: in the float32 instantiation the pair arithmetic is
float32 and every `1.0 / ...`, the spline polynomial and the running sums are float64, as Numba types the
reference source).  The only addition is the row range [i0, i1) (the kernels are independent per i), which is what
allows a bounded sample of the N = 1M workload.

Pinned: tests/test_oracle.py compares it with the C oracle (itself pinned to the reference's golden outputs) and
tools/numba_vs_reference.py — run in the build container, where the reference is importable — shows it is the
same speed as the reference's own functions.

Run as a script it prints one JSON line (bench.py calls it in a subprocess, so a missing or misbehaving Numba
can never take the benchmark down):

    python -m oracle.numba_port --n 1000000 --kernel plummer --precision float32 --seconds 8
"""
from __future__ import annotations

import json
import sys
import time

import numpy as np

try:
    from numba import get_num_threads, njit, prange, set_num_threads
except Exception as e:  # noqa: BLE001 - reported by the caller
    raise ImportError(f"numba is not usable here: {e}") from e

KERNEL_IDS = {"newtonian": 0, "plummer": 1, "dehnen_k1": 2, "dehnen_k2": 3, "spline": 4}


@njit(fastmath=True, inline="always")
def _force_factor(r2, h, kernel_id):
    if kernel_id == 0:
        return 1.0 / (r2 * np.sqrt(r2))
    if kernel_id == 1:
        denom = r2 + h * h
        return 1.0 / (denom * np.sqrt(denom))
    if kernel_id == 2:
        denom = r2 + h * h
        sd = np.sqrt(denom)
        return 1.0 / (denom * sd) + 1.5 * h * h / (denom * denom * sd)
    if kernel_id == 3:
        denom = r2 + h * h
        sd = np.sqrt(denom)
        hh = h * h
        return 1.0 / (denom * sd) + 1.5 * hh / (denom * denom * sd) + 3.75 * hh * hh / (denom * denom * denom * sd)
    r = np.sqrt(r2)
    if r > h:
        return 1.0 / (r * r * r)
    hinv = 1.0 / h
    h3inv = hinv * hinv * hinv
    q = r * hinv
    q2 = q * q
    if kernel_id == 4:
        if q <= 0.5:
            return h3inv * (10.666666666666666 + q2 * (-38.4 + 32.0 * q))
        return h3inv * (21.333333333333333 - 48.0 * q + 38.4 * q2 - 10.666666666666667 * q2 * q
                        - 0.06666666666666667 / (q2 * q))
    return 0.0


@njit(fastmath=True, inline="always")
def _potential_factor(r2, h, kernel_id):
    r = np.sqrt(r2)
    if kernel_id == 0:
        if r > 0:
            return -1.0 / r
        return 0.0
    if kernel_id == 1:
        return -1.0 / np.sqrt(r2 + h * h)
    if kernel_id == 2:
        h2 = h * h
        denom = r2 + h2
        inv_sqrt = 1.0 / np.sqrt(denom)
        return -inv_sqrt - 0.5 * h2 * (inv_sqrt / denom)
    if kernel_id == 3:
        h2 = h * h
        denom = r2 + h2
        sd = np.sqrt(denom)
        inv_sqrt = 1.0 / sd
        inv_d32 = inv_sqrt / denom
        return -inv_sqrt - 0.5 * h2 * inv_d32 - 0.375 * h2 * h2 * (inv_d32 / sd)
    if kernel_id == 4:
        if h == 0.0 or r >= h:
            return -1.0 / r
        hinv = 1.0 / h
        q = r * hinv
        if q < 1e-8:
            return -2.8 * hinv
        q2 = q * q
        if q <= 0.5:
            return (-2.8 + q2 * (5.33333333333333333 + q2 * q2 * (6.4 * q - 9.6))) * hinv
        if q <= 1.0:
            return (-3.2 + 0.066666666666666666666 / q
                    + q2 * (10.666666666666666666666 + q * (-16.0 + q * (9.6 - 2.1333333333333333333333 * q)))) * hinv
        return -1.0 / r
    return 0.0


@njit(parallel=True, fastmath=True)
def _forces_rows(pos, mass, h, kernel_id, r_eps, i0, i1, out):
    n = pos.shape[0]
    for i in prange(i0, i1):
        fx = 0.0
        fy = 0.0
        fz = 0.0
        xi = pos[i, 0]
        yi = pos[i, 1]
        zi = pos[i, 2]
        hi = h[i]
        for j in range(n):
            if i == j:
                continue
            dx = pos[j, 0] - xi
            dy = pos[j, 1] - yi
            dz = pos[j, 2] - zi
            r2 = dx * dx + dy * dy + dz * dz + r_eps * r_eps
            hj = h[j]
            h_ij = hi if hi >= hj else hj
            factor = mass[j] * _force_factor(r2, h_ij, kernel_id)
            fx += factor * dx
            fy += factor * dy
            fz += factor * dz
        out[i - i0, 0] = fx
        out[i - i0, 1] = fy
        out[i - i0, 2] = fz


@njit(parallel=True, fastmath=True)
def _potential_rows(pos, mass, h, kernel_id, r_eps, i0, i1, out):
    n = pos.shape[0]
    eps2 = r_eps * r_eps
    for i in prange(i0, i1):
        xi = pos[i, 0]
        yi = pos[i, 1]
        zi = pos[i, 2]
        hi = h[i]
        pot_i = 0.0
        for j in range(n):
            if i == j:
                continue
            dx = pos[j, 0] - xi
            dy = pos[j, 1] - yi
            dz = pos[j, 2] - zi
            r2 = dx * dx + dy * dy + dz * dz + eps2
            hj = h[j]
            h_eff = hi if hi >= hj else hj
            pot_i += mass[j] * _potential_factor(r2, h_eff, kernel_id)
        out[i - i0] = pot_i


def _operands(pos, mass, softening, precision):
    dt = np.float64 if precision == "float64" else np.float32
    pos = np.ascontiguousarray(pos, dtype=dt)
    n = pos.shape[0]
    mass = np.full(n, mass, dtype=dt) if np.isscalar(mass) else np.ascontiguousarray(mass, dtype=dt)
    h = np.full(n, softening, dtype=dt) if np.isscalar(softening) else np.ascontiguousarray(softening, dtype=dt)
    return pos, mass, h, dt


def compute_nbody_forces_numba(pos, mass, softening=0.0, G=1.0, kernel="spline", precision="float64", rows=None):
    """(rows or all N, 3) accelerations in the precision's dtype (reference compute_nbody_forces_cpu, fields.py:1021-1086)."""
    pos, mass, h, dt = _operands(pos, mass, softening, precision)
    i0, i1 = (0, pos.shape[0]) if rows is None else rows
    out = np.empty((i1 - i0, 3), dtype=dt)
    _forces_rows(pos, mass, h, KERNEL_IDS[kernel], dt(1e-15), i0, i1, out)
    return out * dt(G)


def compute_nbody_potential_numba(pos, mass, softening=0.0, G=1.0, kernel="spline", precision="float64", rows=None):
    """(rows or all N,) potential (reference compute_nbody_potential_cpu, fields.py:1089-1146)."""
    pos, mass, h, dt = _operands(pos, mass, softening, precision)
    i0, i1 = (0, pos.shape[0]) if rows is None else rows
    out = np.empty((i1 - i0,), dtype=dt)
    _potential_rows(pos, mass, h, KERNEL_IDS[kernel], dt(1e-15), i0, i1, out)
    return out * dt(G)


def timed_sample(pos, mass, softening, kernel, precision, seconds, threads=None) -> dict:
    """Force pass + potential pass on rows [0, R) against all N, R sized so that the sample takes about `seconds`
    (after a JIT warm-up that is not timed).  Returns a bench.py cpu_baseline-style dict."""
    if threads:
        set_num_threads(int(threads))
    n = pos.shape[0]
    kw = dict(kernel=kernel, precision=precision)

    def run(rows):
        t0 = time.perf_counter()
        compute_nbody_forces_numba(pos, mass, softening, rows=(0, rows), **kw)
        compute_nbody_potential_numba(pos, mass, softening, rows=(0, rows), **kw)
        return time.perf_counter() - t0

    run(min(n, 64))                                 # JIT compile (seconds), not timed
    rows = min(n, 1024)
    t = run(rows)
    for _ in range(4):                              # the first calls after the JIT run far below the sustained rate
        rate = rows * n / max(t, 1e-9)              # (thread pool, clocks): re-size until the sample takes `seconds`
        rows = int(min(n, max(1024, seconds * rate / n)))
        t = run(rows)
        if t >= 0.7 * seconds or rows >= n:
            break
    t = min(t, run(rows))                           # best of two samples of the final size
    return {"value": rows * n / t, "unit": "interactions/s", "cores": int(get_num_threads()), "kind": "port",
            "technology": "numba @njit(parallel=True, fastmath=True), prange over i (oracle/numba_port.py)",
            "sample": f"rows [0,{rows}) of the N={n} workload against all N, force pass + potential pass ({kernel}, "
                      f"{precision} pair math / float64 sums), {t:.1f} s", "seconds": t, "rows": rows}


def _main():
    import argparse
    import os

    p = argparse.ArgumentParser()
    p.add_argument("--n", type=int, default=100_000)
    p.add_argument("--kernel", default="plummer")
    p.add_argument("--precision", default="float32")
    p.add_argument("--softening", type=float, default=0.01)
    p.add_argument("--seconds", type=float, default=8.0)
    p.add_argument("--threads", type=int, default=0)
    a = p.parse_args()
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from nbody_streams_b200.run import make_plummer_sphere   # host-only generator (no CUDA needed)

    xv, m = make_plummer_sphere(a.n, M_total=1e9, a=1.0, seed=42069)
    prec = "float64" if a.precision == "float64" else "float32"
    res = timed_sample(xv[:, :3], float(m[0]), a.softening, a.kernel, prec, a.seconds, a.threads or None)
    print(json.dumps(res))


if __name__ == "__main__":
    _main()
