"""GPU parity of the PAIR-SYMMETRIC sweeps (csrc/nb_sym_f32.cuh, nb_sym_f64.cuh; large-N default) through the
C ABI: against the C oracle on the same seeded inputs, against the one-sided sweeps, and through
size-independent properties.  `algorithm='symmetric'` forces the path at sizes the oracle finishes in seconds.

Tolerances as in test_gpu_parity.py (BASELINE.json north_star): float64 max_i |da|/|a| <= 1e-12,
float32 / float32_kahan relative RMS <= 1e-5.
"""
import numpy as np
import pytest

import oracle
from conftest import rel_max, rel_rms

pytestmark = pytest.mark.gpu

KERNELS = ["newtonian", "plummer", "dehnen_k1", "dehnen_k2", "spline"]
PRECISIONS = ["float64", "float32", "float32_kahan"]
G = 4.300917270069976e-06


@pytest.fixture(scope="module")
def nb():
    import nbody_streams_b200 as nb

    return nb


def _rounded(a, precision):
    a = np.asarray(a, dtype=np.float64)
    return a if precision == "float64" else a.astype(np.float32).astype(np.float64)


def _check(acc, phi, ref, pref, precision):
    assert np.isfinite(acc).all() and np.isfinite(phi).all()
    if precision == "float64":
        assert rel_max(acc, ref) <= 1e-12
        assert rel_max(phi, pref) <= 1e-12
    else:
        assert rel_rms(acc, ref) <= 1e-5
        assert rel_rms(phi, pref) <= 1e-5


def _oracle(pos, mass, h, kernel, precision, **kw):
    p, m, hh = _rounded(pos, precision), _rounded(mass, precision), _rounded(h, precision)
    ref = oracle.compute_nbody_forces_cpu(p, m, hh, G, kernel, **kw)
    pref = oracle.compute_nbody_potential_cpu(p, m, hh, G, kernel, dehnen_k2_gpu_formula=True, **kw)
    return ref, pref


def _cloud(n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return rng.normal(size=(n, 3)) * scale


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("precision", PRECISIONS)
def test_uniform_species_all_kernels(nb, kernel, precision):
    """One species (scalar mass and softening): every symmetric tile takes the uniform fast path."""
    n = 3000
    pos = _cloud(n, 11)
    acc, phi = nb.compute_nbody_forces_potential_gpu(pos, 2.5e4, 0.05, G, precision=precision, kernel=kernel,
                                                     algorithm="symmetric")
    ref, pref = _oracle(pos, np.full(n, 2.5e4), np.full(n, 0.05), kernel, precision)
    _check(acc, phi, ref, pref, precision)
    # separate force-only / potential-only sweeps give the same numbers
    a1 = nb.compute_nbody_forces_gpu(pos, 2.5e4, 0.05, G, precision=precision, kernel=kernel, algorithm="symmetric")
    p1 = nb.compute_nbody_potential_gpu(pos, 2.5e4, 0.05, G, precision=precision, kernel=kernel,
                                        algorithm="symmetric")
    _check(a1, p1, ref, pref, precision)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("precision", PRECISIONS)
def test_per_particle_mass_and_softening(nb, kernel, precision):
    """Per-particle arrays: every tile takes the general path (max-softening rule and both masses per pair)."""
    n = 2777
    rng = np.random.default_rng(5)
    pos = _cloud(n, 12)
    mass = rng.uniform(1e3, 1e5, n)
    h = rng.uniform(0.01, 0.2, n)
    acc, phi = nb.compute_nbody_forces_potential_gpu(pos, mass, h, G, precision=precision, kernel=kernel,
                                                     algorithm="symmetric")
    ref, pref = _oracle(pos, mass, h, kernel, precision)
    _check(acc, phi, ref, pref, precision)


@pytest.mark.parametrize("kernel", ["plummer", "spline"])
@pytest.mark.parametrize("precision", PRECISIONS)
def test_species_blocks_mix_fast_and_general_tiles(nb, kernel, precision):
    """Three species laid out like run_simulation does (dark, star with per-particle softening, one heavy BH at
    the origin): uniform blocks, a block straddling a species boundary, per-particle tiles, padding."""
    n_d, n_s = 5 * 1024 + 300, 2 * 1024 + 77
    pos = np.vstack([_cloud(n_d, 1), _cloud(n_s, 2, 0.3), np.zeros((1, 3))])
    mass = np.concatenate([np.full(n_d, 1e4), np.full(n_s, 500.0), [1e8]])
    h = np.concatenate([np.full(n_d, 0.05), np.linspace(0.01, 0.03, n_s), [0.001]])
    acc, phi = nb.compute_nbody_forces_potential_gpu(pos, mass, h, G, precision=precision, kernel=kernel,
                                                     algorithm="symmetric")
    ref, pref = _oracle(pos, mass, h, kernel, precision)
    _check(acc, phi, ref, pref, precision)


@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, 1023, 1024, 1025, 2049])
@pytest.mark.parametrize("precision", ["float64", "float32_kahan"])
def test_edge_sizes(nb, n, precision):
    pos = _cloud(n, 100 + n)
    for kernel in ("plummer", "spline"):
        acc, phi = nb.compute_nbody_forces_potential_gpu(pos, 1.0e5, 0.1, G, precision=precision, kernel=kernel,
                                                         algorithm="symmetric")
        assert acc.shape == (n, 3) and phi.shape == (n,)
        if n == 1:
            assert np.all(acc == 0) and np.all(phi == 0)
            continue
        ref, pref = _oracle(pos, np.full(n, 1.0e5), np.full(n, 0.1), kernel, precision)
        _check(acc, phi, ref, pref, precision)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_coincident_particles_and_origin(nb, precision):
    """Pairs at zero separation (softened kernels) and a particle sitting exactly at the origin."""
    n = 1500
    pos = _cloud(n, 9)
    pos[10] = pos[1400]            # coincident pair in different tiles (symmetric tile)
    pos[20] = pos[21]              # coincident pair in the same tile (diagonal tile)
    pos[700] = 0.0
    for kernel in ("plummer", "spline", "dehnen_k1"):
        acc, phi = nb.compute_nbody_forces_potential_gpu(pos, 3.0e4, 0.07, G, precision=precision, kernel=kernel,
                                                         algorithm="symmetric")
        ref, pref = _oracle(pos, np.full(n, 3.0e4), np.full(n, 0.07), kernel, precision)
        _check(acc, phi, ref, pref, precision)


@pytest.mark.parametrize("precision", ["float64", "float32_kahan"])
def test_dense_core_many_pairs_inside_softening(nb, precision):
    """Large softening: most pairs are inside h, so the spline's exact path runs for nearly every step."""
    n = 2100
    pos = _cloud(n, 3, 0.05)
    acc, phi = nb.compute_nbody_forces_potential_gpu(pos, 1.0e4, 0.2, G, precision=precision, kernel="spline",
                                                     algorithm="symmetric")
    ref, pref = _oracle(pos, np.full(n, 1.0e4), np.full(n, 0.2), "spline", precision)
    _check(acc, phi, ref, pref, precision)


@pytest.mark.parametrize("kernel", ["plummer", "spline"])
@pytest.mark.parametrize("precision", PRECISIONS)
def test_many_units_per_cta_against_onesided_and_oracle_rows(nb, kernel, precision):
    """N = 70k: ~9k symmetric work units, several per CTA, segments crossing i-blocks.  Checked against oracle
    rows, against the one-sided sweep on all rows and through Newton's third law."""
    n = 70_001
    xv, m = nb.make_plummer_sphere(n, M_total=1e9, a=1.0, seed=77)
    pos = xv[:, :3]
    kw = dict(precision=precision, kernel=kernel)
    a_s, p_s = nb.compute_nbody_forces_potential_gpu(pos, float(m[0]), 0.02, G, algorithm="symmetric", **kw)
    a_o, p_o = nb.compute_nbody_forces_potential_gpu(pos, float(m[0]), 0.02, G, algorithm="onesided", **kw)
    rows = (34_000, 35_500)
    ref, pref = _oracle(pos, m, np.full(n, 0.02), kernel, precision, rows=rows, fast=True)
    _check(a_s[rows[0]:rows[1]], p_s[rows[0]:rows[1]], ref, pref, precision)
    if precision == "float64":
        assert rel_max(a_s, a_o) <= 1e-12 and rel_max(p_s, p_o) <= 1e-12
    else:
        assert rel_rms(a_s, a_o) <= 2e-6 and rel_rms(p_s, p_o) <= 2e-6
    f = a_s.astype(np.float64) * m[:, None]
    net = np.linalg.norm(f.sum(axis=0)) / (np.median(np.linalg.norm(f, axis=1)) * n)
    assert net < (1e-13 if precision == "float64" else 1000 * 1.2e-7)


def test_auto_selection_threshold(nb):
    from nbody_streams_b200 import _lib
    from nbody_streams_b200.fields import use_symmetric

    lib = _lib.load()
    n0 = lib.nb_sym_min_n()
    assert n0 > 0
    assert use_symmetric(n0) and not use_symmetric(n0 - 1)
    assert not use_symmetric(n0, world=2) and use_symmetric(32768, world=2)     # several ranks: from 32768 on
    assert use_symmetric(10, "symmetric") and not use_symmetric(10**7, "onesided")
    with pytest.raises(ValueError):
        use_symmetric(10, "both")


@pytest.mark.parametrize("precision", ["float64", "float32_kahan"])
def test_n_sources_takes_the_symmetric_path_too(nb, precision):
    """Sources restricted to the first n_sources particles (unbinding): triangle over the source tiles + rectangle of
    the test-particle tiles, against the oracle with the test particles' masses zeroed and against the one-sided sweep."""
    n, n_src = 40_000, 25_001                     # n_src inside a tile
    xv, m = nb.make_plummer_sphere(n, M_total=1e9, a=1.0, seed=21)
    pos = xv[:, :3]
    kw = dict(precision=precision, kernel="spline", n_sources=n_src)
    a_s, p_s = nb.compute_nbody_forces_potential_gpu(pos, m, 0.03, G, algorithm="symmetric", **kw)
    a_o, p_o = nb.compute_nbody_forces_potential_gpu(pos, m, 0.03, G, algorithm="onesided", **kw)
    m0 = m.copy()
    m0[n_src:] = 0.0
    for rows in ((0, 700), (24_700, 25_400), (39_300, 40_000)):      # sources, the straddling tile, test particles
        ref, pref = _oracle(pos, m0, np.full(n, 0.03), "spline", precision, rows=rows, fast=True)
        _check(a_s[rows[0]:rows[1]], p_s[rows[0]:rows[1]], ref, pref, precision)
    if precision == "float64":
        assert rel_max(a_s, a_o) <= 1e-12 and rel_max(p_s, p_o) <= 1e-12
    else:
        assert rel_rms(a_s, a_o) <= 2e-6 and rel_rms(p_s, p_o) <= 2e-6
    # n_sources = 0: nothing gravitates
    a0, p0 = nb.compute_nbody_forces_potential_gpu(pos, m, 0.03, G, precision=precision, kernel="spline", n_sources=0,
                                                   algorithm="symmetric")
    assert np.all(a0 == 0) and np.all(p0 == 0)
