"""BASELINE.json configurations at FULL size on one GPU: parity on an oracle row subset (the reference
kernels are independent per i, so rows [i0,i1) against all N are exact) plus size-independent properties
(Newton's third law, energy bookkeeping).  SURVEY.md 8(d) gives the concrete inputs."""
import numpy as np
import pytest

import oracle
from conftest import rel_max, rel_rms

pytestmark = pytest.mark.gpu
G = 4.300917270069976e-06


@pytest.fixture(scope="module")
def nb():
    import nbody_streams_b200 as nb

    return nb


def _rounded(a):
    return a.astype(np.float32).astype(np.float64)


def _newton3(acc, mass):
    f = acc.astype(np.float64) * np.asarray(mass, dtype=np.float64)[:, None]
    return np.linalg.norm(f.sum(axis=0)) / (np.median(np.linalg.norm(f, axis=1)) * len(f))


def test_c2_two_species_120k_float32(nb):
    """C2: dark 100k (M=1e9, a=1, h=0.05) + star 20k (M=1e7, a=0.3, h=linspace(0.01,0.03)), float32, spline."""
    xd, md = nb.make_plummer_sphere(100_000, M_total=1e9, a=1.0, seed=42069)
    xs, ms = nb.make_plummer_sphere(20_000, M_total=1e7, a=0.3, seed=42070)
    pos = np.vstack([xd[:, :3], xs[:, :3]])
    sp = [nb.Species.dark(100_000, 1e4, softening=0.05),
          nb.Species.star(20_000, 500.0, softening=np.linspace(0.01, 0.03, 20_000))]
    from nbody_streams_b200.species import _build_particle_arrays

    mass, h = _build_particle_arrays(sp)
    for kernel in ("spline", "plummer"):
        acc, phi = nb.compute_nbody_forces_potential_gpu(pos, mass, h, precision="float32", kernel=kernel)
        rows = (99_000, 101_024)                      # straddles the species boundary
        kw = dict(kernel=kernel, rows=rows, fast=True)
        ref = oracle.compute_nbody_forces_cpu(_rounded(pos), _rounded(mass), _rounded(h), **kw)
        pref = oracle.compute_nbody_potential_cpu(_rounded(pos), _rounded(mass), _rounded(h), **kw)
        assert rel_rms(acc[rows[0]:rows[1]], ref) <= 1e-5
        assert rel_rms(phi[rows[0]:rows[1]], pref) <= 1e-5
        # also against the reference's own float32 CPU path semantics (float pair math, double sums)
        ref32 = oracle.compute_nbody_forces_cpu(pos, mass, h, precision="float32", **kw)
        assert rel_rms(acc[rows[0]:rows[1]], ref32) <= 1e-5
        assert _newton3(acc, mass) < 1000 * 1.2e-7       # reference fields.py __main__ criterion: < 1000 eps


def test_c3_plummer_1m_float32_kahan_force_and_potential(nb):
    """C3 (the bench workload): Plummer N=1M, scalar softening 0.01, float32_kahan, force + potential."""
    xv, m = nb.make_plummer_sphere(1_000_000, M_total=1e9, a=1.0, seed=42069)
    pos = xv[:, :3]
    rows = (500_000, 502_048)
    for kernel in ("plummer", "spline"):
        acc, phi = nb.compute_nbody_forces_potential_gpu(pos, float(m[0]), 0.01, precision="float32_kahan",
                                                         kernel=kernel)
        kw = dict(kernel=kernel, rows=rows, fast=True)
        ref = oracle.compute_nbody_forces_cpu(_rounded(pos), float(np.float32(m[0])), float(np.float32(0.01)), **kw)
        pref = oracle.compute_nbody_potential_cpu(_rounded(pos), float(np.float32(m[0])), float(np.float32(0.01)), **kw)
        assert rel_rms(acc[rows[0]:rows[1]], ref) <= 1e-5
        assert rel_rms(phi[rows[0]:rows[1]], pref) <= 1e-5
        assert _newton3(acc, np.full(len(acc), m[0])) < 1000 * 1.2e-7
        # separate force-only / potential-only entry points agree with the fused sweep
        a2 = nb.compute_nbody_forces_gpu(pos, float(m[0]), 0.01, precision="float32_kahan", kernel=kernel)
        assert rel_rms(a2, acc) <= 1e-6
    # virial sanity of the generated system: 2 KE / |PE| ~ 1 with PE from the GPU potential
    ke = 0.5 * np.sum(m * np.sum(xv[:, 3:] ** 2, axis=1))
    pe = 0.5 * np.sum(m * phi.astype(np.float64))
    assert abs(ke / abs(pe) - 0.5) < 0.01


def test_c4_plummer_262k_float64(nb, tmp_path):
    """C4: Plummer N=262144, softening 0.02, float64: row-subset parity at 1e-12 and a short KDK run."""
    N = 262_144
    xv, m = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=42069)
    rows = (131_000, 132_024)
    for kernel in ("spline", "plummer"):
        acc, phi = nb.compute_nbody_forces_potential_gpu(xv[:, :3], m, 0.02, precision="float64", kernel=kernel)
        ref = oracle.compute_nbody_forces_cpu(xv[:, :3], m, 0.02, kernel=kernel, rows=rows)
        pref = oracle.compute_nbody_potential_cpu(xv[:, :3], m, 0.02, kernel=kernel, rows=rows)
        assert rel_max(acc[rows[0]:rows[1]], ref) <= 1e-12
        assert rel_max(phi[rows[0]:rows[1]], pref) <= 1e-12
        assert _newton3(acc, m) < 1000 * 2.2e-16
    diags = []
    nb.run_nbody_gpu(xv, m, 0.0, 10 * 1e-4, 1e-4, 0.02, precision="float64", kernel="plummer",
                     output_dir=str(tmp_path), save_snapshots=False, verbose=False, diagnostics=diags,
                     diagnostics_every=5)
    E = np.array([d[2] + d[3] for d in diags])
    assert np.abs(E - E[0]).max() / abs(E[0]) < 1e-6
    p = np.array([d[4:7] for d in diags])
    assert np.abs(p - p[0]).max() / (1e9 * np.sqrt(2 * diags[0][2] / 1e9)) < 1e-12


def test_c5_stream_progenitor_2m_float32(nb):
    """C5: dark 1.6M (h=0.05) + star 400k (h=0.02) + one 1e8 Msun BH (h=0.001) at the origin, float32, spline."""
    xd, md = nb.make_plummer_sphere(1_600_000, M_total=1e9, a=1.0, seed=1)
    xs, ms = nb.make_plummer_sphere(400_000, M_total=2e7, a=0.3, seed=2)
    pos = np.vstack([xd[:, :3], xs[:, :3], np.zeros((1, 3))])
    mass = np.concatenate([md, ms, [1e8]])
    h = np.concatenate([np.full(1_600_000, 0.05), np.full(400_000, 0.02), [0.001]])
    acc, phi = nb.compute_nbody_forces_potential_gpu(pos, mass, h, precision="float32", kernel="spline")
    for rows in ((1_599_500, 1_600_524), (1_999_000, 2_000_001)):    # species boundary; stars + the BH row
        kw = dict(kernel="spline", rows=rows, fast=True)
        ref = oracle.compute_nbody_forces_cpu(_rounded(pos), _rounded(mass), _rounded(h), **kw)
        pref = oracle.compute_nbody_potential_cpu(_rounded(pos), _rounded(mass), _rounded(h), **kw)
        assert rel_rms(acc[rows[0]:rows[1]], ref) <= 1e-5
        assert rel_rms(phi[rows[0]:rows[1]], pref) <= 1e-5
    assert np.isfinite(acc).all() and np.isfinite(phi).all()
    assert _newton3(acc, mass) < 1000 * 1.2e-7
