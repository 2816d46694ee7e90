"""Iterative unbinding: the oracle restatement against the reference's own output
(tests/golden/unbinding.npz, made by tests/golden/make_golden_unbinding.py), and the host-side checks of
nbody_streams_b200.unbind that need no GPU."""
from pathlib import Path

import numpy as np
import pytest

import oracle
from nbody_streams_b200 import unbind

GOLD = np.load(Path(__file__).parent / "golden" / "unbinding.npz")


def test_oracle_case_a_two_components_history():
    xd, xs = GOLD["a_xd"], GOLD["a_xs"]
    (bd, bs, hd, hs), cp, cv = oracle.iterative_unbinding_cpu(
        xd[:, :3], xd[:, 3:], GOLD["a_md"], xs[:, :3], xs[:, 3:], GOLD["a_ms"], softening=0.03, center_on="dark",
        vel_aperture=5.0, return_history=True)
    np.testing.assert_array_equal(bd, GOLD["a_bound_dark"])
    np.testing.assert_array_equal(bs, GOLD["a_bound_star"])
    np.testing.assert_array_equal(np.array(hd), GOLD["a_hist_dark"])
    np.testing.assert_array_equal(np.array(hs), GOLD["a_hist_star"])
    np.testing.assert_allclose(cp, GOLD["a_cpos"], rtol=1e-13)
    np.testing.assert_allclose(cv, GOLD["a_cvel"], rtol=1e-13)
    assert bd.dtype.kind == "i" and 0 < bd.sum() < len(bd)


def test_oracle_case_b_center_on_both_scalar_masses():
    xd, xs = GOLD["a_xd"], GOLD["a_xs"]
    (bd, bs), cp, cv = oracle.iterative_unbinding_cpu(
        xd[:, :3], xd[:, 3:], float(GOLD["a_md"][0]), xs[:, :3], xs[:, 3:], float(GOLD["a_ms"][0]), softening=0.05,
        center_on="both", vel_aperture=0.3, top_fraction=0.05)
    np.testing.assert_array_equal(bd, GOLD["b_bound_dark"])
    np.testing.assert_array_equal(bs, GOLD["b_bound_star"])
    np.testing.assert_allclose(cp, GOLD["b_cpos"], rtol=1e-13)
    np.testing.assert_allclose(cv, GOLD["b_cvel"], rtol=1e-13)


def test_oracle_case_c_given_centre_iteration_cap():
    xd = GOLD["c_xd"]
    (bd,), cp, cv = oracle.iterative_unbinding_cpu(
        xd[:, :3], xd[:, 3:], GOLD["c_md"], center_position=[20.0, -5.0, 3.0], center_velocity=[100.0, 50.0, -30.0],
        recursive_iter_converg=3, softening=0.03)
    np.testing.assert_array_equal(bd, GOLD["c_bound_dark"])
    np.testing.assert_array_equal(cp, GOLD["c_cpos"])
    np.testing.assert_array_equal(cv, GOLD["c_cvel"])


# ---- host-side behaviour of the product function (no GPU needed: all of these raise before any device work)

def _args(n=16):
    rng = np.random.default_rng(0)
    return rng.standard_normal((n, 3)), rng.standard_normal((n, 3)), 1.0


@pytest.mark.parametrize("method", ["tree", "tree_gpu", "bfe", "direct"])
def test_other_solvers_are_named_and_refused(method):
    with pytest.raises(ValueError, match="direct_gpu"):
        unbind.iterative_unbinding(*_args(), potential_compute_method=method)


def test_default_method_is_the_references_and_unknown_is_rejected():
    import inspect

    sig = inspect.signature(unbind.iterative_unbinding)
    names = list(sig.parameters)
    assert names[:17] == ["pos_dark", "vel_dark", "mass_dark", "pos_star", "vel_star", "mass_star", "center_position",
                          "center_velocity", "recursive_iter_converg", "potential_compute_method", "softening", "G",
                          "center_on", "vel_aperture", "tol_frac_change", "verbose", "return_history"]
    assert sig.parameters["potential_compute_method"].default == "tree"      # reference utils/main.py:1732
    assert sig.parameters["softening"].default == 0.03 and sig.parameters["recursive_iter_converg"].default == 50
    for extra in ("precision", "kernel", "device", "group"):
        assert sig.parameters[extra].kind is inspect.Parameter.KEYWORD_ONLY
    with pytest.raises(ValueError, match="Unknown"):
        unbind.iterative_unbinding(*_args(), potential_compute_method="fmm")


def test_input_validation_errors():
    p, v, m = _args()
    with pytest.raises(ValueError, match="shape"):
        unbind.iterative_unbinding(p[:, :2], v, m, potential_compute_method="direct_gpu")
    with pytest.raises(ValueError, match="shape"):
        unbind.iterative_unbinding(p, v[:5], m, potential_compute_method="direct_gpu")
    with pytest.raises(ValueError, match="mass length"):
        unbind.iterative_unbinding(p, v, np.ones(3), potential_compute_method="direct_gpu")
    with pytest.raises(ValueError, match="requires star data"):
        unbind.iterative_unbinding(p, v, m, potential_compute_method="direct_gpu", center_on="star")
    with pytest.raises(ValueError, match="Precision"):
        unbind.iterative_unbinding(p, v, m, potential_compute_method="direct_gpu", precision="float16")
    with pytest.raises(ValueError, match="Invalid kernel"):
        unbind.iterative_unbinding(p, v, m, potential_compute_method="direct_gpu", kernel="cubic")


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ImportError):
        unbind.iterative_unbinding(*_args(), potential_compute_method="direct_gpu", verbose=False)
