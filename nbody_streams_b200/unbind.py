"""Iterative unbinding on the GPU — the ``potential_compute_method='direct_gpu'`` leg of the
reference's ``utils.iterative_unbinding`` (reference utils/main.py:1722-2079), a caller of the
direct-summation potential.

Same signature, defaults, return value and iteration as the reference:

    centre  : phi of the full system, centroid of the lowest-phi ``top_fraction`` of the
              ``center_on`` particles, velocity = mass-weighted mean inside ``vel_aperture``
              (utils/main.py:1933-1985)
    iterate : phi_i = potential of the currently bound particles at every particle,
              bound_i = phi_i + v_i^2/2 < 0, until the mask changes by < tol_frac_change
              (utils/main.py:1998-2066)

What is different underneath.  The reference multiplies the masses by the bound mask and runs the
full N x N sweep every iteration with a host round trip of pos, mass and phi
(utils/main.py:2034-2037).  Here the snapshot is uploaded once; each iteration is a stable
partition (bound particles first), one pack and one sweep of all N targets against only the tiles
holding bound sources (N x N_bound pair work), all on the device; the only per-iteration host
traffic is the two scalars (N_bound, changed fraction).  With a torch.distributed ``group`` (one
rank per GPU, every rank holding the snapshot) the target rows are sharded over the ranks and the
phi rows all-gathered — no other exchange.

Only ``'direct_gpu'`` is provided: the tree / BFE / CPU solvers of the reference are outside this
library (ValueError names them).  There is no CPU fallback.
"""
from __future__ import annotations

import logging

import numpy as np

from . import _lib
from .fields import EPS2_DEFAULT, G_DEFAULT, KERNEL_MAP, _PRECISION_MAP, _run_direct
from .species import _emit_performance_warnings

__all__ = ["iterative_unbinding", "compute_iterative_boundness"]

_OTHER_METHODS = ("tree", "tree_gpu", "bfe", "direct")


def _positions(pos, label):
    pos = np.asarray(pos, dtype=float)
    if pos.ndim != 2 or pos.shape[-1] != 3:
        # the reference's validator also lets (N,) radii through (utils/_validation.py:52-56);
        # they cannot be unbound against, so they are rejected here
        raise ValueError(f"{label} must have shape (N, 3), got {pos.shape}")
    return pos


def _velocities(vel, n, label):
    vel = np.asarray(vel, dtype=float)
    if vel.ndim != 2 or vel.shape != (n, 3):
        raise ValueError(f"{label} must have shape ({n}, 3), got {vel.shape}")
    return vel


def _masses(mass, n):
    # reference utils/_validation.py:78-116
    if mass is None:
        return np.ones(n, dtype=float)
    mass = np.asarray(mass, dtype=float)
    if mass.ndim == 0:
        return np.full(n, float(mass), dtype=float)
    if mass.shape[0] != n:
        raise ValueError(f"mass length ({mass.shape[0]}) does not match number of particles ({n})")
    return mass


def iterative_unbinding(
    pos_dark,
    vel_dark,
    mass_dark,
    pos_star=None,
    vel_star=None,
    mass_star=None,
    center_position=[],
    center_velocity=[],
    recursive_iter_converg: int = 50,
    potential_compute_method: str = "tree",
    softening: float = 0.03,
    G: float = G_DEFAULT,
    center_on: str = "dark",
    vel_aperture: float = 5.0,
    tol_frac_change: float = 0.0001,
    verbose: bool = True,
    return_history: bool = False,
    *,
    precision: str = "float32_kahan",
    kernel: str = "spline",
    device=None,
    group=None,
    **kwargs,
) -> tuple:
    """Bound/unbound masks by iterative unbinding; drop-in for reference
    ``iterative_unbinding(..., potential_compute_method='direct_gpu')`` (utils/main.py:1722).

    Returns ``((bound_dark[, bound_star][, history_dark[, history_star]]), center_position,
    center_velocity)`` with integer masks (1 = bound), as the reference does.

    Keyword-only extensions (defaults = what the reference's direct_gpu leg computes,
    ``compute_nbody_potential_gpu`` defaults, fields.py:866-875): ``precision``, ``kernel``,
    ``device``, ``group``.  ``top_fraction`` is read from ``**kwargs`` like the reference
    (utils/main.py:1902); ``theta`` / ``lmax`` are accepted and unused.
    """
    method = potential_compute_method.lower()
    if method != "direct_gpu":
        if method in _OTHER_METHODS:
            raise ValueError(
                f"potential_compute_method='{potential_compute_method}' is not part of nbody_streams_b200; "
                "this library provides the GPU direct-summation leg only: use 'direct_gpu'."
            )
        raise ValueError(f"Unknown potential_compute_method '{potential_compute_method}'")
    if precision.lower() not in _PRECISION_MAP:
        raise ValueError(f"Precision must be one of {list(_PRECISION_MAP.keys())}")
    if kernel.lower() not in KERNEL_MAP:
        raise ValueError(f"Invalid kernel '{kernel}'. Must be one of: {list(KERNEL_MAP.keys())}")

    pos_dark = _positions(pos_dark, "pos_dark")
    n_dark = pos_dark.shape[0]
    vel_dark = _velocities(vel_dark, n_dark, "vel_dark")
    mass_dark_arr = _masses(mass_dark, n_dark)
    has_stars = pos_star is not None
    n_star = 0
    if has_stars:
        pos_star = _positions(pos_star, "pos_star")
        n_star = pos_star.shape[0]
        vel_star = _velocities(vel_star, n_star, "vel_star")
        mass_star_arr = _masses(mass_star, n_star)
    if center_on == "star" and not has_stars:
        raise ValueError("center_on='star' requires star data.")

    torch = _lib.require_cuda()
    from .engine import group_info

    n_total = n_dark + n_star
    _emit_performance_warnings(n_total, "gpu", "direct")

    logger = logging.getLogger("boundness")
    if verbose and not logger.hasHandlers():
        handler = logging.StreamHandler()
        handler.setLevel(logging.INFO)
        handler.setFormatter(logging.Formatter("[%(levelname)s] %(message)s"))
        logger.addHandler(handler)
    logger.setLevel(logging.INFO if verbose else logging.WARNING)

    top_fraction = kwargs.get("top_fraction", 0.01)
    prec, dtype_np = _PRECISION_MAP[precision.lower()]
    kernel_id = KERNEL_MAP[kernel.lower()]
    eps2 = EPS2_DEFAULT[precision.lower()]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    rank, world = group_info(group) if group is not None else (0, 1)

    def up(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(device)

    # whole snapshot resident on the device, float64 like the reference's host arrays
    if has_stars:
        pos_all = torch.cat((up(pos_dark), up(pos_star)))
        vel_all = torch.cat((up(vel_dark), up(vel_star)))
        mass_all = torch.cat((up(mass_dark_arr), up(mass_star_arr)))
    else:
        pos_all, vel_all, mass_all = up(pos_dark), up(vel_dark), up(mass_dark_arr)

    def potential(pos_t, mass_t, n_sources=None):
        """phi (float64, device) at every particle from the first n_sources particles."""
        _, phi = _run_direct(pos_t, mass_t, 0.0, None, float(softening), n_total, prec, kernel_id, _lib.MODE_POT,
                             G, eps2, dtype_np, group if world > 1 else None, n_sources)
        return phi.double()

    if center_on == "both" or not has_stars:
        ctr_sl = slice(None)
    elif center_on == "star":
        ctr_sl = slice(n_dark, None)
    else:
        ctr_sl = slice(None, n_dark)

    center_position = np.asarray(center_position, dtype=float)
    center_velocity = np.asarray(center_velocity, dtype=float)

    with torch.cuda.device(device):
        if center_position.size < 3:
            phi_ctr = potential(pos_all, mass_all)[ctr_sl]
            n_pick = max(10, int(phi_ctr.shape[0] * top_fraction))
            idxs = torch.argsort(phi_ctr, stable=True)[:n_pick]
            w = mass_all[ctr_sl][idxs]
            center_position = ((pos_all[ctr_sl][idxs] * w[:, None]).sum(0) / w.sum()).cpu().numpy()
        cpos_t = up(center_position)

        if center_velocity.size < 3:
            pos_ctr, vel_ctr, mass_ctr = pos_all[ctr_sl], vel_all[ctr_sl], mass_all[ctr_sl]
            dist2 = ((pos_ctr - cpos_t) ** 2).sum(1)
            sel = dist2 < vel_aperture ** 2
            if not bool(sel.any()):
                sel = torch.ones_like(sel)
            w = mass_ctr * sel
            center_velocity = ((vel_ctr * w[:, None]).sum(0) / w.sum()).cpu().numpy()
        cvel_t = up(center_velocity)

        logger.info("Center position (density_peak): %s", np.around(center_position, 2))
        logger.info("Center velocity: %s", np.around(center_velocity, 2))

        pos_rel = pos_all - cpos_t
        kin = 0.5 * ((vel_all - cvel_t) ** 2).sum(1)      # does not change between iterations

        mask = torch.ones(n_total, dtype=torch.bool, device=device)
        n_bound = n_total
        history = []
        min_particles = 5

        for it in range(recursive_iter_converg):
            if n_bound < min_particles:
                logger.info("Stopping: only %d particles remaining.", n_bound)
                break
            if n_bound == n_total:
                phi = potential(pos_rel, mass_all)
            else:
                # stable partition: bound particles first, so that only their tiles are swept as sources
                order = torch.argsort(~mask, stable=True)
                phi_p = potential(pos_rel[order].contiguous(), (mass_all * mask)[order].contiguous(), n_bound)
                phi = torch.empty_like(phi_p)
                phi[order] = phi_p
            bound = (phi + kin) < 0
            if return_history:
                history.append(bound.clone())
            changed = bound != mask
            # one small D2H per iteration: the two scalars the host loop branches on
            frac_change, n_bound = (float(x) for x in torch.stack((changed.double().mean(), bound.sum().double())).tolist())
            n_bound = int(n_bound)
            logger.info("Iter %d: delta bound mask = %.5f", it, frac_change)
            mask = bound
            if frac_change < tol_frac_change:
                logger.info("Converged after %d iterations.", it + 1)
                break

        mask_np = mask.cpu().numpy()
        results: list = [mask_np[:n_dark].astype(int)]
        if has_stars:
            results.append(mask_np[n_dark:].astype(int))
        if return_history:
            hist = [h.cpu().numpy() for h in history]
            results.append([h[:n_dark].copy() for h in hist])
            if has_stars:
                results.append([h[n_dark:].copy() for h in hist])
    return tuple(results), center_position, center_velocity


def compute_iterative_boundness(*args, **kwargs):
    """Deprecated alias kept by the reference (utils/main.py:1714-1720)."""
    import warnings

    warnings.warn("compute_iterative_boundness is deprecated; use iterative_unbinding.", DeprecationWarning, stacklevel=2)
    return iterative_unbinding(*args, **kwargs)
