"""oracle — CPU restatement of the reference's direct-summation path.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this
package; nothing under ``nbody_streams_b200/`` does.

Public functions mirror the reference's CPU wrappers (same argument names and
defaults, reference nbody_streams/fields.py:1021-1146):

    compute_nbody_forces_cpu(pos, mass, softening, G, kernel, nthreads, precision)
    compute_nbody_potential_cpu(...)

plus ``rows=(i0, i1)`` to evaluate a row subset against all N (exact, because
the reference kernels are independent per i) and ``fast=True`` to use the
-O3/-ffast-math/AVX build for timing.  Parity is pinned by tests/test_oracle.py
against golden vectors produced by the reference itself
(tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent

G_DEFAULT = 4.300917270069976e-06  # reference fields.py:90
KERNEL_MAP = {"newtonian": 0, "plummer": 1, "dehnen_k1": 2, "dehnen_k2": 3, "spline": 4}  # fields.py:94-100
_PREC = {"float64": np.float64, "float32": np.float32, "float32_kahan": np.float32}  # fields.py:104-115

_libs: dict = {}


def build(force: bool = False) -> None:
    """Compile the two shared libraries with oracle/Makefile (gcc + OpenMP)."""
    targets = [_HERE / "libnbody_oracle.so", _HERE / "libnbody_oracle_fast.so"]
    src = _HERE / "nbody_oracle.c"
    if not force and all(t.exists() and t.stat().st_mtime >= src.stat().st_mtime for t in targets):
        return
    subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, capture_output=True)


def _lib(fast: bool):
    key = bool(fast)
    if key not in _libs:
        path = _HERE / ("libnbody_oracle_fast.so" if fast else "libnbody_oracle.so")
        if not path.exists():
            build()
        lib = C.CDLL(str(path))
        sfx = "_fast" if fast else ""
        dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
        getattr(lib, "nbo_forces_f64" + sfx).argtypes = [dp, dp, dp, C.c_int, C.c_double, C.c_long, C.c_long, C.c_long, dp]
        getattr(lib, "nbo_forces_f32" + sfx).argtypes = [fp, fp, fp, C.c_int, C.c_float, C.c_long, C.c_long, C.c_long, fp]
        getattr(lib, "nbo_potential_f64" + sfx).argtypes = [dp, dp, dp, C.c_int, C.c_double, C.c_long, C.c_long, C.c_long, C.c_int, dp]
        getattr(lib, "nbo_potential_f32" + sfx).argtypes = [fp, fp, fp, C.c_int, C.c_float, C.c_long, C.c_long, C.c_long, C.c_int, fp]
        getattr(lib, "nbo_num_threads" + sfx).restype = C.c_int
        for n in ("nbo_forces_f64", "nbo_forces_f32", "nbo_potential_f64", "nbo_potential_f32"):
            getattr(lib, n + sfx).restype = None
        _libs[key] = (lib, sfx)
    return _libs[key]


def num_threads(fast: bool = False) -> int:
    lib, sfx = _lib(fast)
    return int(getattr(lib, "nbo_num_threads" + sfx)())


def set_num_threads(n: int, fast: bool = False) -> None:
    """Override OMP_NUM_THREADS (torchrun exports 1) for the timing legs."""
    lib, sfx = _lib(fast)
    getattr(lib, "nbo_set_num_threads" + sfx)(int(n))


def _prepare(pos, mass, softening, precision, kernel):
    """Host-side validation + scalar broadcast, as reference fields.py:485-565 and :618-658."""
    key = precision.lower()
    if key not in _PREC:
        raise ValueError(f"Precision must be one of {list(_PREC.keys())}")
    dt = _PREC[key]
    pos = np.ascontiguousarray(np.asarray(pos, dtype=dt))
    if pos.ndim != 2 or pos.shape[1] != 3:
        raise ValueError(f"pos must have shape (N, 3), got {pos.shape}")
    n = pos.shape[0]

    def one(v, label):
        if np.isscalar(v):
            return np.full(n, float(v), dtype=dt)
        a = np.ascontiguousarray(np.asarray(v, dtype=dt))
        if a.shape[0] != n:
            raise ValueError(f"{label} must have length N={n}, got {a.shape[0]}")
        return a

    m = one(mass, "mass")
    h = one(softening, "softening")
    if kernel.lower() not in KERNEL_MAP:
        raise ValueError(f"Invalid kernel '{kernel}'. Must be one of: {list(KERNEL_MAP.keys())}")
    return pos, m, h, n, dt, KERNEL_MAP[kernel.lower()]


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def compute_nbody_forces_cpu(pos, mass, softening=0.0, G=G_DEFAULT, kernel="spline", nthreads=None,
                             precision="float64", *, rows=None, fast=False):
    """Restates reference compute_nbody_forces_cpu (fields.py:1021-1086): G * sum_j m_j K dx,
    r_eps = dtype(1e-15) added squared to r^2, max-softening rule, self skipped."""
    pos, m, h, n, dt, kid = _prepare(pos, mass, softening, precision, kernel)
    i0, i1 = (0, n) if rows is None else rows
    lib, sfx = _lib(fast)
    out = np.zeros((i1 - i0, 3), dtype=dt)
    if dt == np.float64:
        getattr(lib, "nbo_forces_f64" + sfx)(_ptr(pos, C.c_double), _ptr(m, C.c_double), _ptr(h, C.c_double), kid,
                                               float(dt(1e-15)), n, i0, i1, _ptr(out, C.c_double))
    else:
        getattr(lib, "nbo_forces_f32" + sfx)(_ptr(pos, C.c_float), _ptr(m, C.c_float), _ptr(h, C.c_float), kid,
                                               float(dt(1e-15)), n, i0, i1, _ptr(out, C.c_float))
    return G * out   # float32 array * Python float stays float32, as in the reference


def compute_nbody_potential_cpu(pos, mass, softening=0.0, G=G_DEFAULT, kernel="spline", nthreads=None,
                                precision="float64", *, rows=None, fast=False, dehnen_k2_gpu_formula=False):
    """Restates reference compute_nbody_potential_cpu (fields.py:1088-1146).

    ``dehnen_k2_gpu_formula=True`` evaluates the dehnen_k2 potential the way the reference
    GPU kernel does (cuda_kernels.py:394-402) instead of the CPU line fields.py:287."""
    pos, m, h, n, dt, kid = _prepare(pos, mass, softening, precision, kernel)
    i0, i1 = (0, n) if rows is None else rows
    lib, sfx = _lib(fast)
    out = np.zeros(i1 - i0, dtype=dt)
    if dt == np.float64:
        getattr(lib, "nbo_potential_f64" + sfx)(_ptr(pos, C.c_double), _ptr(m, C.c_double), _ptr(h, C.c_double), kid,
                                                  float(dt(1e-15)), n, i0, i1, int(dehnen_k2_gpu_formula),
                                                  _ptr(out, C.c_double))
    else:
        getattr(lib, "nbo_potential_f32" + sfx)(_ptr(pos, C.c_float), _ptr(m, C.c_float), _ptr(h, C.c_float), kid,
                                                  float(dt(1e-15)), n, i0, i1, int(dehnen_k2_gpu_formula),
                                                  _ptr(out, C.c_float))
    return G * out


def run_nbody_cpu_direct(phase_space, masses, time_start, time_end, dt, softening, G=G_DEFAULT, kernel="spline",
                         precision="float64", fast=False, record_every=0):
    """KDK leapfrog of reference run_nbody_cpu, direct method, no hooks / I/O
    (run.py:1110-1147): kick(dt/2) -> drift(dt) -> t += dt -> forces -> kick(dt/2).

    Returns (final (N,6) float64, diagnostics list of (step, KE, PE, px, py, pz) when record_every > 0)."""
    xv = np.array(phase_space, dtype=np.float64, copy=True)
    masses = np.asarray(masses, dtype=np.float64)
    total_steps = int(round((time_end - time_start) / dt))   # run.py:1022
    acc = compute_nbody_forces_cpu(xv[:, :3], masses, softening, G, kernel, precision=precision, fast=fast)
    diags = []

    def diag(step):
        phi = compute_nbody_potential_cpu(xv[:, :3], masses, softening, G, kernel, precision=precision, fast=fast)
        v2 = np.sum(xv[:, 3:6] ** 2, axis=1)
        ke = 0.5 * np.dot(masses, v2)           # run.py:1069-1073
        pe = 0.5 * np.dot(masses, phi)
        p = masses @ xv[:, 3:6]
        diags.append((step, float(ke), float(pe), float(p[0]), float(p[1]), float(p[2])))

    if record_every:
        diag(0)
    for step in range(1, total_steps + 1):
        xv[:, 3:6] += acc * (dt / 2)
        xv[:, 0:3] += xv[:, 3:6] * dt
        acc = compute_nbody_forces_cpu(xv[:, :3], masses, softening, G, kernel, precision=precision, fast=fast)
        xv[:, 3:6] += acc * (dt / 2)
        if record_every and (step % record_every == 0 or step == total_steps):
            diag(step)
    return xv, diags


def iterative_unbinding_cpu(pos_dark, vel_dark, mass_dark, pos_star=None, vel_star=None, mass_star=None,
                            center_position=(), center_velocity=(), recursive_iter_converg=50, softening=0.03,
                            G=G_DEFAULT, center_on="dark", vel_aperture=5.0, tol_frac_change=0.0001,
                            return_history=False, top_fraction=0.01, precision="float64", kernel="spline"):
    """Restatement of the reference's iterative_unbinding with potential_compute_method='direct'
    (reference utils/main.py:1722-2079; centre :1933-1985, loop :1998-2066, output :2068-2079).
    phi comes from this oracle's compute_nbody_potential_cpu (reference default precision float64,
    kernel 'spline', fields.py:1096-1104).  Pinned against the reference's own output in
    tests/golden/unbinding_*.npz."""
    pos_dark = np.asarray(pos_dark, dtype=float)
    vel_dark = np.asarray(vel_dark, dtype=float)
    n_dark = pos_dark.shape[0]
    bc = lambda m, n: np.full(n, float(m)) if np.ndim(m) == 0 else np.asarray(m, dtype=float)
    mass_dark = bc(mass_dark, n_dark)
    has_stars = pos_star is not None
    if has_stars:
        pos_star = np.asarray(pos_star, dtype=float)
        vel_star = np.asarray(vel_star, dtype=float)
        mass_star = bc(mass_star, pos_star.shape[0])
        pos_all = np.vstack((pos_dark, pos_star))
        vel_all = np.vstack((vel_dark, vel_star))
        mass_all = np.concatenate((mass_dark, mass_star))
    else:
        pos_all, vel_all, mass_all = pos_dark.copy(), vel_dark.copy(), mass_dark.copy()
    if center_on == "both" or not has_stars:
        sl = slice(None)
    elif center_on == "star":
        sl = slice(n_dark, None)
    else:
        sl = slice(None, n_dark)

    def pot(p, m):
        return compute_nbody_potential_cpu(p, m, softening=softening, G=G, kernel=kernel, precision=precision)

    center_position = np.asarray(center_position, dtype=float)
    center_velocity = np.asarray(center_velocity, dtype=float)
    if center_position.size < 3:                                    # main.py:1936-1972
        phi_c = pot(pos_all, mass_all)[sl]
        n_pick = max(10, int(len(phi_c) * top_fraction))
        idx = np.argsort(phi_c)[:n_pick]
        center_position = np.average(pos_all[sl][idx], axis=0, weights=mass_all[sl][idx])
    if center_velocity.size < 3:                                    # main.py:1974-1985
        d2 = np.sum((pos_all[sl] - center_position) ** 2, axis=1)
        sel = d2 < vel_aperture ** 2
        if np.any(sel):
            center_velocity = np.average(vel_all[sl][sel], axis=0, weights=mass_all[sl][sel])
        else:
            center_velocity = np.average(vel_all[sl], axis=0, weights=mass_all[sl])
    pos_rel = pos_all - center_position
    vel_rel = vel_all - center_velocity
    mask = np.ones(len(pos_all), dtype=bool)
    hist = []
    energies = None
    for _ in range(recursive_iter_converg):                         # main.py:1998-2066
        if int(mask.sum()) < 5:
            break
        phi = pot(pos_rel, mass_all * mask)
        energies = phi + 0.5 * np.sum(vel_rel ** 2, axis=1)
        bound = energies < 0
        hist.append(bound.copy())
        frac = np.mean(bound != mask)
        mask = bound
        if frac < tol_frac_change:
            break
    res = [mask[:n_dark].astype(int)]
    if has_stars:
        res.append(mask[n_dark:].astype(int))
    if return_history:
        res.append([h[:n_dark] for h in hist])
        if has_stars:
            res.append([h[n_dark:] for h in hist])
    return tuple(res), center_position, center_velocity
