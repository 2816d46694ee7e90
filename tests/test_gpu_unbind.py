"""GPU iterative unbinding against the reference's own output (tests/golden/unbinding.npz)
and, at larger N, the oracle restatement; plus the N x n_sources potential sweep it is built on."""
from pathlib import Path

import numpy as np
import pytest

import oracle
from nbody_streams_b200 import (compute_nbody_forces_potential_gpu, compute_nbody_potential_gpu, make_plummer_sphere)
from nbody_streams_b200.unbind import compute_iterative_boundness, iterative_unbinding

pytestmark = pytest.mark.gpu

GOLD = np.load(Path(__file__).parent / "golden" / "unbinding.npz")


def _case_a(**kw):
    xd, xs = GOLD["a_xd"], GOLD["a_xs"]
    return iterative_unbinding(xd[:, :3], xd[:, 3:], GOLD["a_md"], xs[:, :3], xs[:, 3:], GOLD["a_ms"],
                               potential_compute_method="direct_gpu", softening=0.03, center_on="dark",
                               vel_aperture=5.0, verbose=False, **kw)


def test_float64_reproduces_the_reference_masks_and_history():
    (bd, bs, hd, hs), cp, cv = _case_a(return_history=True, precision="float64")
    np.testing.assert_array_equal(bd, GOLD["a_bound_dark"])
    np.testing.assert_array_equal(bs, GOLD["a_bound_star"])
    assert len(hd) == len(GOLD["a_hist_dark"]) and len(hs) == len(GOLD["a_hist_star"])
    np.testing.assert_array_equal(np.array(hd), GOLD["a_hist_dark"])
    np.testing.assert_array_equal(np.array(hs), GOLD["a_hist_star"])
    np.testing.assert_allclose(cp, GOLD["a_cpos"], rtol=1e-12)
    np.testing.assert_allclose(cv, GOLD["a_cvel"], rtol=1e-12)
    assert bd.dtype == GOLD["a_bound_dark"].dtype and bd.shape == (3000,) and bs.shape == (600,)


def test_default_precision_matches_up_to_marginal_particles():
    """float32_kahan (what the reference's direct_gpu leg runs): phi differs by ~1e-7 relative, so only
    particles within that of E = 0 may flip."""
    (bd, bs), cp, cv = _case_a()
    flips = int((bd != GOLD["a_bound_dark"]).sum() + (bs != GOLD["a_bound_star"]).sum())
    assert flips <= 2, flips
    np.testing.assert_allclose(cp, GOLD["a_cpos"], rtol=1e-5)
    np.testing.assert_allclose(cv, GOLD["a_cvel"], rtol=1e-5)


def test_center_on_both_scalar_masses_small_aperture():
    xd, xs = GOLD["a_xd"], GOLD["a_xs"]
    (bd, bs), cp, cv = iterative_unbinding(
        xd[:, :3], xd[:, 3:], float(GOLD["a_md"][0]), xs[:, :3], xs[:, 3:], float(GOLD["a_ms"][0]),
        potential_compute_method="direct_gpu", softening=0.05, center_on="both", vel_aperture=0.3, verbose=False,
        top_fraction=0.05, precision="float64")
    np.testing.assert_array_equal(bd, GOLD["b_bound_dark"])
    np.testing.assert_array_equal(bs, GOLD["b_bound_star"])
    np.testing.assert_allclose(cp, GOLD["b_cpos"], rtol=1e-12)
    np.testing.assert_allclose(cv, GOLD["b_cvel"], rtol=1e-12)


def test_given_centre_iteration_cap_ragged_n_and_deprecated_alias():
    xd = GOLD["c_xd"]
    kw = dict(center_position=[20.0, -5.0, 3.0], center_velocity=[100.0, 50.0, -30.0], recursive_iter_converg=3,
              potential_compute_method="direct_gpu", softening=0.03, verbose=False, precision="float64")
    (bd,), cp, cv = iterative_unbinding(xd[:, :3], xd[:, 3:], GOLD["c_md"], **kw)
    np.testing.assert_array_equal(bd, GOLD["c_bound_dark"])
    np.testing.assert_array_equal(cp, GOLD["c_cpos"])
    np.testing.assert_array_equal(cv, GOLD["c_cvel"])
    with pytest.warns(DeprecationWarning):
        (bd2,), _, _ = compute_iterative_boundness(xd[:, :3], xd[:, 3:], GOLD["c_md"], **kw)
    np.testing.assert_array_equal(bd2, bd)


def _stripped_satellite(n, seed):
    rng = np.random.default_rng(seed)
    xv, m = make_plummer_sphere(n, M_total=1e9, a=1.0, seed=seed)
    pick = rng.random(n) < 0.3
    xv[pick, :3] *= rng.uniform(1.0, 8.0, pick.sum())[:, None]
    xv[pick, 3:] *= rng.uniform(0.8, 4.0, pick.sum())[:, None]
    xv += np.array([-3.0, 8.0, 1.0, -40.0, 10.0, 220.0])
    return xv, m


@pytest.mark.parametrize("kernel", ["spline", "plummer"])
def test_larger_system_against_the_oracle(kernel):
    xv, m = _stripped_satellite(20_011, 5)
    (ref,), cpr, cvr = oracle.iterative_unbinding_cpu(xv[:, :3], xv[:, 3:], m, softening=0.02, kernel=kernel)
    (bd,), cp, cv = iterative_unbinding(xv[:, :3], xv[:, 3:], m, potential_compute_method="direct_gpu", softening=0.02,
                                        verbose=False, precision="float64", kernel=kernel)
    assert 0.5 * len(ref) < ref.sum() < 0.9 * len(ref)
    np.testing.assert_array_equal(bd, ref)
    np.testing.assert_allclose(cp, cpr, rtol=1e-12)
    np.testing.assert_allclose(cv, cvr, rtol=1e-12)
    (bk,), _, _ = iterative_unbinding(xv[:, :3], xv[:, 3:], m, potential_compute_method="direct_gpu", softening=0.02,
                                      verbose=False, kernel=kernel)
    assert (bk != ref).sum() <= 3


def test_everything_unbound_stops_below_five_particles():
    rng = np.random.default_rng(1)
    pos = rng.standard_normal((300, 3))
    vel = 1e4 * rng.standard_normal((300, 3))       # far above the escape speed of 300 solar masses
    (bd,), _, _ = iterative_unbinding(pos, vel, 1.0, center_position=[0, 0, 0], center_velocity=[0, 0, 0],
                                      potential_compute_method="direct_gpu", verbose=False)
    assert bd.sum() == 0


# ---- the N x n_sources sweep

@pytest.mark.parametrize("precision,tol", [("float64", 1e-12), ("float32_kahan", 2e-6), ("float32", 1e-5)])
@pytest.mark.parametrize("n,n_src", [(3000, 1000), (3000, 256), (3000, 1), (3000, 0), (3000, 3000), (1001, 513)])
def test_potential_from_the_first_n_sources_particles(precision, tol, n, n_src):
    rng = np.random.default_rng(n + n_src)
    pos = rng.standard_normal((n, 3))
    mass = rng.uniform(0.5, 2.0, n)
    h = rng.uniform(0.01, 0.1, n)
    got = compute_nbody_potential_gpu(pos, mass, h, precision=precision, kernel="spline", n_sources=n_src)
    m0 = mass.copy()
    m0[n_src:] = 0.0
    dt = np.float64 if precision == "float64" else np.float32
    ref = oracle.compute_nbody_potential_cpu(pos.astype(dt).astype(np.float64), m0.astype(dt).astype(np.float64),
                                             h.astype(dt).astype(np.float64), kernel="spline", precision="float64")
    if n_src == 0:
        assert np.all(got == 0)
        return
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() / scale < tol


def test_n_sources_scalar_mass_forces_and_validation():
    rng = np.random.default_rng(3)
    pos = rng.standard_normal((2000, 3))
    acc, phi = compute_nbody_forces_potential_gpu(pos, 2.0, 0.05, precision="float64", kernel="plummer", n_sources=700)
    m0 = np.full(2000, 2.0)
    m0[700:] = 0
    np.testing.assert_allclose(acc, oracle.compute_nbody_forces_cpu(pos, m0, 0.05, kernel="plummer"), rtol=1e-10,
                               atol=1e-12 * np.abs(acc).max())
    np.testing.assert_allclose(phi, oracle.compute_nbody_potential_cpu(pos, m0, 0.05, kernel="plummer"), rtol=1e-12)
    with pytest.raises(ValueError, match="n_sources"):
        compute_nbody_potential_gpu(pos, 2.0, 0.05, n_sources=2001)


def test_unbinding_above_the_symmetric_threshold_against_the_oracle():
    """N = 40k: the potential sweeps run the pair-symmetric kernels (sources = bound particles)."""
    xv, m = _stripped_satellite(40_000, 11)
    (ref,), cpr, cvr = oracle.iterative_unbinding_cpu(xv[:, :3], xv[:, 3:], m, softening=0.02, kernel="spline")
    (bd,), cp, cv = iterative_unbinding(xv[:, :3], xv[:, 3:], m, potential_compute_method="direct_gpu", softening=0.02,
                                        verbose=False, precision="float64", kernel="spline")
    assert 0.5 * len(ref) < ref.sum() < 0.9 * len(ref)
    assert (bd != ref).sum() <= 1          # double atomics: a particle with |E| within 1e-15 |phi| of 0 may flip
    np.testing.assert_allclose(cp, cpr, rtol=1e-10)
    np.testing.assert_allclose(cv, cvr, rtol=1e-10)
