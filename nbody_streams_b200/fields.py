"""GPU force / potential dispatch — drop-in for the GPU half of the reference's
``nbody_streams/fields.py``.

Same public names, argument order, defaults, return dtypes and error types as
the reference:

* ``compute_nbody_forces_gpu``    (reference fields.py:664-864)
* ``compute_nbody_potential_gpu`` (reference fields.py:866-1018)
* ``get_gpu_info``                (reference fields.py:1149-1185)
* ``G_DEFAULT``, ``KERNEL_MAP``   (reference fields.py:90-100)

What changes underneath: instead of CuPy-JIT-compiled kernel strings
(reference fields.py:417-483) the calls go through ctypes into the compiled
sm_100a library ``libnbody_sm100.so``; device arrays are torch CUDA tensors
(no CuPy in this stack).  ``return_cupy=True`` returns a :class:`DeviceArray`
(a ``torch.Tensor`` subclass with the ``.get()`` / ``.astype()`` /
``__cuda_array_interface__`` surface reference callers use).

There is no CPU fallback in this module.
"""
from __future__ import annotations

import ctypes as C
from typing import Literal

import numpy as np

from . import _lib
from ._lib import DT_F32, DT_F64, MODE_BOTH, MODE_FORCE, MODE_POT, PREC_F32, PREC_F32_COMP, PREC_F64

# Default gravitational constant (kpc, km/s, Msun) — reference fields.py:90
G_DEFAULT = 4.300917270069976e-06

KERNEL_TYPES = Literal["newtonian", "plummer", "dehnen_k1", "dehnen_k2", "spline"]

# reference fields.py:94-100
KERNEL_MAP = {
    "newtonian": 0,
    "plummer": 1,
    "dehnen_k1": 2,
    "dehnen_k2": 3,
    "spline": 4,
}

# precision string -> (library precision code, numpy dtype); reference fields.py:104-115
_PRECISION_MAP = {
    "float64": (PREC_F64, np.float64),
    "float32": (PREC_F32, np.float32),
    "float32_kahan": (PREC_F32_COMP, np.float32),
}

# Regularisation added to r^2.  The reference GPU path adds 1e-15 (fields.py:825),
# its CPU path adds (1e-15)^2 = 1e-30 (fields.py:237,1083-1086).  float64 follows the
# CPU convention (needed for 1e-12 agreement with the Numba path), float32 the GPU one.
EPS2_DEFAULT = {"float64": 1e-30, "float32": 1e-15, "float32_kahan": 1e-15}


def _torch():
    return _lib.require_cuda()


class DeviceArray:  # replaced below by a torch.Tensor subclass once torch is importable
    pass


def _make_device_array_type():
    import torch

    class DeviceArray(torch.Tensor):
        """torch CUDA tensor with the small CuPy-like surface reference callers rely on
        (``.get()`` -> NumPy, reference _chandrasekhar.py:48-54 and run.py:204;
        ``.astype``; ``__cuda_array_interface__`` comes from torch.Tensor)."""

        def get(self) -> np.ndarray:
            return self.detach().as_subclass(torch.Tensor).cpu().numpy()

        def astype(self, dtype, copy: bool = True):
            tdt = _np_to_torch_dtype(np.dtype(dtype))
            out = self.to(tdt)
            if copy and out.data_ptr() == self.data_ptr():
                out = out.clone()
            return out.as_subclass(DeviceArray)

        def __array__(self, dtype=None, copy=None):
            a = self.get()
            return a.astype(dtype) if dtype is not None else a

    return DeviceArray


try:  # torch is part of the stack; keep module importable for the CPU-only unit tests
    DeviceArray = _make_device_array_type()  # type: ignore[misc]
except Exception:  # pragma: no cover
    pass


def _np_to_torch_dtype(dt: np.dtype):
    import torch

    return {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[np.dtype(dt)]


def _is_device_array(x) -> bool:
    """True for torch CUDA tensors and anything exposing __cuda_array_interface__ (CuPy...)."""
    try:
        import torch

        if isinstance(x, torch.Tensor):
            return x.is_cuda
    except ImportError:  # pragma: no cover
        pass
    return hasattr(x, "__cuda_array_interface__")


def _as_cuda_tensor(x, device):
    import torch

    if isinstance(x, torch.Tensor):
        return x.as_subclass(torch.Tensor)
    return torch.as_tensor(x, device=device)  # consumes __cuda_array_interface__ zero-copy


# ---------------------------------------------------------------------------
# validation (host side; identical checks / messages to reference fields.py:485-565)
# ---------------------------------------------------------------------------
def _validate_inputs(pos, mass, softening, precision, kernel):
    prec_key = precision.lower()
    if prec_key not in _PRECISION_MAP:
        raise ValueError(f"Precision must be one of {list(_PRECISION_MAP.keys())}")
    dtype_np = _PRECISION_MAP[prec_key][1]

    if _is_device_array(pos):
        if pos.ndim != 2 or pos.shape[1] != 3:
            raise ValueError(f"pos must have shape (N, 3), got {tuple(pos.shape)}")
        N = pos.shape[0]
        pos_cpu = None
    else:
        pos_cpu = np.asarray(pos, dtype=dtype_np)
        if pos_cpu.ndim != 2 or pos_cpu.shape[1] != 3:
            raise ValueError(f"pos must have shape (N, 3), got {pos_cpu.shape}")
        N = pos_cpu.shape[0]

    def _one(val, label):
        if np.isscalar(val):
            return None, float(val)
        if _is_device_array(val):
            if val.shape[0] != N:
                raise ValueError(f"{label} must have length N={N}, got {val.shape[0]}")
            return None, None
        arr = np.asarray(val, dtype=dtype_np)
        if arr.ndim == 0:
            return None, float(arr)
        if arr.shape[0] != N:
            raise ValueError(f"{label} must have length N={N}, got {arr.shape[0]}")
        return arr, None

    mass_cpu, mass_scalar = _one(mass, "mass")
    h_cpu, h_scalar = _one(softening, "softening")

    if kernel.lower() not in KERNEL_MAP:
        raise ValueError(f"Invalid kernel '{kernel}'. Must be one of: {list(KERNEL_MAP.keys())}")

    return {
        "N": N,
        "pos_cpu": pos_cpu,
        "mass_cpu": mass_cpu,
        "mass_scalar": mass_scalar,
        "h_cpu": h_cpu,
        "h_scalar": h_scalar,
        "kernel_id": KERNEL_MAP[kernel.lower()],
        "dtype_np": dtype_np,
        "prec": _PRECISION_MAP[prec_key][0],
        "original_pos": pos,
        "original_mass": mass,
        "original_h": softening,
    }


# ---------------------------------------------------------------------------
# device-side plumbing
# ---------------------------------------------------------------------------
_GRID_CACHE: dict = {}
_SCRATCH: dict = {}


def _grid_for(prec: int, kernel_id: int, mode: int, device, sym: bool = False) -> int:
    lib = _lib.load()
    key = (device.index, prec, kernel_id, mode, "sym" if sym else lib.nb_get_variant(prec))
    g = _GRID_CACHE.get(key)
    if g is None:
        out = C.c_int(0)
        query = lib.nb_sym_grid if sym else lib.nb_sweep_grid
        _lib.check(query(prec, kernel_id, mode, C.byref(out)))
        g = _GRID_CACHE[key] = int(out.value)
    return g


SYM_MIN_N_MULTI_RANK = 32768


def use_symmetric(N: int, algorithm: str | None = None, world: int = 1) -> bool:
    """Which sweep evaluates an all-pairs problem of N particles: the pair-symmetric one (every unordered
    pair once, applied to both particles; large N) or the one-sided one (every ordered pair, bitwise
    reproducible, finer work units; small N).  `algorithm`: None/'auto', 'symmetric' or 'onesided'.
    'auto' switches at nb_sym_min_n() particles (8192) on one rank and at 32768 on several ranks, where the
    symmetric path adds a reduce-scatter and each rank's launches are 1/world of the work."""
    if algorithm in (None, "auto"):
        n0 = _lib.load().nb_sym_min_n()
        return N >= (n0 if world == 1 else max(n0, SYM_MIN_N_MULTI_RANK))
    if algorithm not in ("symmetric", "onesided"):
        raise ValueError(f"algorithm must be 'auto', 'symmetric' or 'onesided', got {algorithm!r}")
    return algorithm == "symmetric"


def _scratch(device, tag: str, nbytes: int):
    """Grow-only scratch (packed particle records, sweep workspace) per (device, CUDA stream): kernels that use
    a buffer are ordered by the stream they run on, so calls issued on different streams (e.g. from a
    force_extra hook running on run_nbody_gpu's hook stream) never share one.  Not thread-safe, like the
    reference's module-level kernel cache (fields.py:396)."""
    import torch

    key = (device.index, tag, torch.cuda.current_stream(device).cuda_stream)
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = _SCRATCH[key] = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
    return buf


def _dt_code(t) -> int:
    import torch

    if t.dtype == torch.float64:
        return DT_F64
    if t.dtype == torch.float32:
        return DT_F32
    raise ValueError(f"unsupported dtype {t.dtype}")


def _device_operand(arr_cpu, scalar, original, dtype_t, device, rows=slice(None)):
    """-> (tensor or None, scalar): host array -> H2D, device array -> as is, scalar -> broadcast in the pack kernel
    (`rows`: only these rows of an array operand)."""
    import torch

    if scalar is not None:
        return None, float(scalar)
    if arr_cpu is not None:
        return torch.from_numpy(np.ascontiguousarray(arr_cpu[rows])).to(device), 0.0
    t = _as_cuda_tensor(original, device)[rows]
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(dtype_t)
    return t.contiguous(), 0.0


_PINNED_SIZES_SEEN: set = set()


def _to_host(t):
    """Device tensor -> NumPy through page-locked memory: one asynchronous copy at PCIe speed instead of the
    staged pageable path of ``Tensor.cpu()`` (the array owns its pinned block, which returns to torch's caching
    host allocator when the array dies).  Falls back to ``.cpu()`` where pinning is unavailable."""
    import torch

    t = t.as_subclass(torch.Tensor)
    if t.numel() == 0:
        return t.cpu().numpy()
    try:
        # Page-locking fresh memory costs ~1 ms/MB (cudaHostAlloc), far more than the copy itself, and the caller
        # typically still holds the previous result when the next one is produced.  The first request of a size
        # therefore pins three blocks and hands two straight back to torch's caching host allocator: later calls
        # find a cached block as long as no more than two earlier results of that size are alive.
        key = (t.dtype, tuple(t.shape))
        nbytes = t.numel() * t.element_size()
        # (bounded: results between 1 MB and 256 MB, at most 8 distinct shapes per process — torch's host allocator
        # never returns pinned blocks to the OS)
        if key not in _PINNED_SIZES_SEEN and (1 << 20) <= nbytes <= (256 << 20) and len(_PINNED_SIZES_SEEN) < 8:
            _PINNED_SIZES_SEEN.add(key)
            spare = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for _ in range(2)]
            del spare
        out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    except RuntimeError:
        return t.cpu().numpy()
    out.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return out.numpy()


def _run_direct(pos_t, mass_t, mass_scalar, h_t, h_scalar, N, prec, kernel_id, mode, G, eps2, dtype_np, group=None,
                n_sources=None, algorithm=None, gather="all", local_rows=False):
    """pack + sweep + finish on the current stream.  Returns (acc|None, phi|None) torch tensors.

    With a torch.distributed `group` of more than one rank every rank packs all N particles and evaluates its
    share of the pair work: with the pair-symmetric sweep 1/world of the unordered pairs (ring schedule,
    engine.sym_plan) followed by a reduce-scatter of the partial sums, with the one-sided sweep the i-rows of
    its tile-aligned shard.  `gather`: 'all' (default) — the rows are all-gathered and every rank returns the
    full arrays; 'root' — only rank 0 of the group receives them (the others return their own rows);
    'local' — no gather, every rank returns the rows of its shard (``ShardLayout(N, world).rows(rank)``).

    `local_rows`: pos_t / mass_t / h_t hold only this rank's rows; each rank packs its own shard and the packed
    shards are all-gathered (otherwise every rank packs all N from full arrays).

    `n_sources` (None = N): only the j-tiles holding particles [0, n_sources) are swept, i.e. the first
    n_sources particles are the sources and all N are targets.  The caller passes zero masses for the
    particles >= n_sources (the tile straddling n_sources is swept whole)."""
    import torch

    from .engine import (ShardLayout, allgather_shards, gather_rows_padded, group_info, reduce_scatter_shards,
                         sym_plan)

    if gather not in ("all", "root", "local"):
        raise ValueError(f"gather must be 'all', 'root' or 'local', got {gather!r}")
    lib = _lib.load()
    device = pos_t.device
    dtype_t = _np_to_torch_dtype(dtype_np)
    rank, world = group_info(group) if group is not None else (0, 1)
    layout = ShardLayout(N, world)
    i_begin, n_i = layout.first(rank), layout.count(rank)
    # multi-rank: outputs are allocated shard-sized (padded) so the all-gather needs no staging copy
    n_rows = layout.shard if world > 1 else n_i
    acc_buf = torch.empty((n_rows, 3), dtype=dtype_t, device=device) if mode & MODE_FORCE else None
    phi_buf = torch.empty((n_rows,), dtype=dtype_t, device=device) if mode & MODE_POT else None
    acc = acc_buf[:n_i] if acc_buf is not None else None
    phi = phi_buf[:n_i] if phi_buf is not None else None
    if N == 0:
        return acc, phi
    sym = n_sources is None and use_symmetric(N, algorithm, world)
    # sources restricted to the first n_sources particles on ONE rank: still pair-symmetric — the pairs inside the
    # source set once (triangle over the source tiles), and (test-particle tiles) x (source tiles) as a rectangle whose
    # j-side contributions vanish with the test particles' zero masses
    sym_sources = n_sources is not None and world == 1 and use_symmetric(N, algorithm)
    with torch.cuda.device(device):
        grid = _grid_for(prec, kernel_id, mode, device, sym=sym)
        n_slots = layout.n_slots if world > 1 else lib.nb_padded_count(N)
        packed = _scratch(device, "packed", lib.nb_packed_bytes(prec, n_slots))
        stream = torch.cuda.current_stream(device).cuda_stream
        mass_ptr, mass_dt = (mass_t.data_ptr(), _dt_code(mass_t)) if mass_t is not None else (None, 0)
        h_ptr, h_dt = (h_t.data_ptr(), _dt_code(h_t)) if h_t is not None else (None, 0)
        acc_ptr = acc_buf.data_ptr() if acc_buf is not None else None
        phi_ptr = phi_buf.data_ptr() if phi_buf is not None else None
        if world == 1 and sym:
            ws_bytes = lib.nb_sym_workspace_bytes(n_slots)
            ws = _scratch(device, "workspace", ws_bytes)
            _lib.check(lib.nb_direct_sym(prec, kernel_id, mode, pos_t.data_ptr(), _dt_code(pos_t), mass_ptr, mass_dt,
                                         float(mass_scalar), h_ptr, h_dt, float(h_scalar), N, float(G), float(eps2),
                                         packed.data_ptr(), ws.data_ptr(), ws_bytes, grid, acc_ptr, phi_ptr, stream))
        elif sym_sources:
            grid = _grid_for(prec, kernel_id, mode, device, sym=True)
            tile = lib.nb_tile_j()
            n_src = max(0, min(int(n_sources), N))
            t_src, t_all = lib.nb_padded_count(n_src) // tile, n_slots // tile      # source tiles, all tiles
            ws_bytes = lib.nb_sym_workspace_bytes(n_slots)
            ws = _scratch(device, "workspace", ws_bytes)
            _lib.check(lib.nb_pack(prec, kernel_id, packed.data_ptr(), 0, pos_t.data_ptr(), _dt_code(pos_t),
                                   mass_ptr, mass_dt, float(mass_scalar), h_ptr, h_dt, float(h_scalar), N, n_slots,
                                   float(eps2), stream))
            args = (float(eps2), ws.data_ptr(), ws_bytes, n_slots)
            if t_src > 0:       # pairs inside the source tiles (the straddling tile carries zero-mass test particles)
                _lib.check(lib.nb_sym_sweep(prec, kernel_id, mode, packed.data_ptr(), N, 0, t_src, 1, 0, 0, 0, 0, 0, 0,
                                            0, 0, *args, 1, grid, stream))
            else:
                ws[:ws_bytes].zero_()
            if t_all > t_src and t_src > 0:   # remaining (test-particle) tiles against the source tiles
                _lib.check(lib.nb_sym_sweep(prec, kernel_id, mode, packed.data_ptr(), N, t_src, t_all, 0, 0, t_src, 0, 0,
                                            0, 0, 0, 0, *args, 0, grid, stream))
            _lib.check(lib.nb_sym_finish(prec, mode, ws.data_ptr(), 0, N, float(G), acc_ptr, phi_ptr, stream))
        elif world == 1 and n_sources is None:
            ws_bytes = lib.nb_workspace_bytes(prec, N, grid, 1)
            ws = _scratch(device, "workspace", ws_bytes)
            _lib.check(lib.nb_direct(prec, kernel_id, mode, pos_t.data_ptr(), _dt_code(pos_t), mass_ptr, mass_dt,
                                     float(mass_scalar), h_ptr, h_dt, float(h_scalar), N, float(G), float(eps2),
                                     packed.data_ptr(), ws.data_ptr(), ws_bytes, grid, acc_ptr, phi_ptr, stream))
        else:
            if local_rows:
                # own shard only (records [i_begin, i_begin + shard), padding included), then the exchange
                _lib.check(lib.nb_pack(prec, kernel_id, packed.data_ptr(), i_begin, pos_t.data_ptr(), _dt_code(pos_t),
                                       mass_ptr, mass_dt, float(mass_scalar), h_ptr, h_dt, float(h_scalar), n_i,
                                       layout.shard, float(eps2), stream))
                allgather_shards(packed, layout, rank, int(lib.nb_packed_bytes(prec, layout.shard)), group)
            else:
                _lib.check(lib.nb_pack(prec, kernel_id, packed.data_ptr(), 0, pos_t.data_ptr(), _dt_code(pos_t),
                                       mass_ptr, mass_dt, float(mass_scalar), h_ptr, h_dt, float(h_scalar), N, n_slots,
                                       float(eps2), stream))
            if sym:
                ws_bytes = lib.nb_sym_workspace_bytes(n_slots)
                ws = _scratch(device, "workspace", ws_bytes)
                ws_rows = ws[:ws_bytes].view(torch.float64).view(n_slots, 4)
                plan = sym_plan(layout, rank, lib.nb_sym_block_tiles(prec))
                own = plan["own"]
                if own[1] > own[0]:
                    A, B, Cr, ibC = plan["A"], plan["B"], plan["C"], plan["ibC"]
                    _lib.check(lib.nb_sym_sweep(prec, kernel_id, mode, packed.data_ptr(), N, own[0], own[1], 1,
                                                A[0], A[1], B[0], B[1], Cr[0], Cr[1], ibC[0], ibC[1], float(eps2),
                                                ws.data_ptr(), ws_bytes, n_slots, 1, grid, stream))
                else:
                    ws_rows.zero_()
                mine = torch.empty((layout.shard, 4), dtype=torch.float64, device=device)
                reduce_scatter_shards(ws_rows, mine, layout, rank, group)
                if n_i > 0:
                    _lib.check(lib.nb_sym_finish(prec, mode, mine.data_ptr(), 0, n_i, float(G), acc_ptr, phi_ptr,
                                                 stream))
            elif n_i > 0:
                ws_bytes = lib.nb_workspace_bytes(prec, n_i, grid, 1)
                ws = _scratch(device, "workspace", ws_bytes)
                n_src = N if n_sources is None else max(0, min(int(n_sources), N))
                n_tiles = C.c_int(lib.nb_padded_count(n_src) // lib.nb_tile_j())
                _lib.check(lib.nb_sweep(prec, kernel_id, mode, packed.data_ptr(), N, 0, n_tiles.value, 0, 0,
                                        i_begin, n_i, float(eps2), ws.data_ptr(), ws_bytes, 0, grid, stream))
                _lib.check(lib.nb_finish(prec, kernel_id, mode, n_i, grid, 1, C.byref(n_tiles), ws.data_ptr(),
                                         float(G), acc_ptr, phi_ptr, stream))
            if world > 1 and gather != "local":
                if acc_buf is not None:
                    acc = gather_rows_padded(acc_buf, layout, rank, group, root_only=(gather == "root"), local=acc)
                if phi_buf is not None:
                    phi = gather_rows_padded(phi_buf, layout, rank, group, root_only=(gather == "root"), local=phi)
    return acc, phi


def _limit_sources(mass_t, mass_s, N, n_sources, dtype_np, device):
    """Masses with zeros from particle n_sources on (test particles), as a device tensor."""
    import torch

    if n_sources is None:
        return mass_t, mass_s, None
    n_src = int(n_sources)
    if not 0 <= n_src <= N:
        raise ValueError(f"n_sources must be in [0, N={N}], got {n_sources}")
    if mass_t is None:
        m = torch.full((N,), float(mass_s), dtype=_np_to_torch_dtype(dtype_np), device=device)
    else:
        m = mass_t.clone()
    m[n_src:] = 0
    return m, 0.0, n_src


def _prepare(pos, mass, softening, precision, kernel, skip_validation, device, group=None, n_sources=None):
    """Common front end of the public calls -> device operands (+ `local`: the operands hold only the rows of
    this rank's shard).  With a multi-rank `group` each rank uploads / keeps only its own rows: the packed
    records of the other shards arrive through the all-gather over NVLink instead of world x redundant PCIe
    copies (not with n_sources, whose mass masking needs the full arrays)."""
    import torch

    from .engine import ShardLayout, group_info

    rank, world = group_info(group) if group is not None else (0, 1)
    local = world > 1 and n_sources is None

    def rows_of(N):
        return ShardLayout(N, world).rows(rank)

    if skip_validation:
        # Hot-loop path (reference fields.py:790-799): trust device arrays of the right dtype.
        prec_key = precision.lower()
        prec, dtype_np = _PRECISION_MAP[prec_key]
        kernel_id = KERNEL_MAP[kernel.lower()]
        dev = device or (pos.device if hasattr(pos, "device") and isinstance(pos, torch.Tensor) else torch.device("cuda", torch.cuda.current_device()))
        pos_t = _as_cuda_tensor(pos, dev)
        N = pos_t.shape[0]
        mass_t, mass_s = (None, float(mass)) if np.isscalar(mass) else (_as_cuda_tensor(mass, dev), 0.0)
        h_t, h_s = (None, float(softening)) if np.isscalar(softening) else (_as_cuda_tensor(softening, dev), 0.0)
        if local:
            r = rows_of(N)
            pos_t = pos_t[r]
            mass_t = mass_t[r] if mass_t is not None else None
            h_t = h_t[r] if h_t is not None else None
        cont = lambda t: t.contiguous() if t is not None else None  # noqa: E731
        return pos_t.contiguous(), cont(mass_t), mass_s, cont(h_t), h_s, N, prec, kernel_id, dtype_np, local

    v = _validate_inputs(pos, mass, softening, precision, kernel)
    dtype_np = v["dtype_np"]
    dtype_t = _np_to_torch_dtype(dtype_np)
    if device is None:
        if v["pos_cpu"] is None and isinstance(pos, torch.Tensor):
            device = pos.device
        else:
            device = torch.device("cuda", torch.cuda.current_device())
    r = rows_of(v["N"]) if local else slice(None)
    if v["pos_cpu"] is None:
        pos_t = _as_cuda_tensor(v["original_pos"], device)[r]
        if pos_t.dtype not in (torch.float32, torch.float64):
            pos_t = pos_t.to(dtype_t)
        pos_t = pos_t.contiguous()
    else:
        pos_t = torch.from_numpy(np.ascontiguousarray(v["pos_cpu"][r])).to(device)
    mass_t, mass_s = _device_operand(v["mass_cpu"], v["mass_scalar"], v["original_mass"], dtype_t, device, r)
    h_t, h_s = _device_operand(v["h_cpu"], v["h_scalar"], v["original_h"], dtype_t, device, r)
    return pos_t, mass_t, mass_s, h_t, h_s, v["N"], v["prec"], v["kernel_id"], dtype_np, local


# ---------------------------------------------------------------------------
# PUBLIC API
# ---------------------------------------------------------------------------
def compute_nbody_forces_gpu(
    pos,
    mass,
    softening=0.0,
    G: float = G_DEFAULT,
    precision: Literal["float32", "float64", "float32_kahan"] = "float32_kahan",
    kernel: KERNEL_TYPES = "spline",
    return_cupy: bool = False,
    skip_validation: bool = False,
    *,
    eps2: float | None = None,
    device=None,
    group=None,
    n_sources=None,
    algorithm=None,
    gather="all",
):
    """Direct-summation gravitational accelerations on the GPU.

    Drop-in for reference ``compute_nbody_forces_gpu`` (fields.py:664-864):
    ``a_i = G sum_{j != i} m_j K(r_ij^2 + eps2, max(h_i, h_j)) (r_j - r_i)``.

    Parameters are those of the reference.  ``pos`` (N,3), ``mass`` scalar or
    (N,), ``softening`` scalar or (N,) may be NumPy arrays, torch tensors (CPU
    or CUDA) or any object exposing ``__cuda_array_interface__``.  Returns an
    (N,3) array of the kernel dtype (float32 for 'float32'/'float32_kahan',
    float64 for 'float64'): NumPy by default, a device array when
    ``return_cupy=True``.

    Keyword-only extensions (defaults reproduce the reference call): ``eps2``
    overrides the r^2 regularisation (see ``EPS2_DEFAULT``), ``device`` picks
    the CUDA device, ``group`` (a torch.distributed process group, one rank per
    GPU) shards the i-particles over its ranks and all-gathers the result,
    ``n_sources`` makes only the first n_sources particles gravitating sources
    (the others are massless test particles and their source tiles are never
    swept: N x n_sources pair work, used by ``unbind.iterative_unbinding``),
    ``algorithm`` ('auto' | 'symmetric' | 'onesided', see :func:`use_symmetric`), ``gather`` (with
    ``group``: 'all' — every rank returns all N rows (default); 'root' — only rank 0 does, the others return
    the rows of their shard; 'local' — every rank returns the rows of its shard, no gather).

    Raises ImportError without a CUDA device / built library, ValueError for
    bad shapes, precision or kernel names, RuntimeError for CUDA failures.
    """
    _torch()
    ops = _prepare(pos, mass, softening, precision, kernel, skip_validation, device, group, n_sources)
    pos_t, mass_t, mass_s, h_t, h_s, N, prec, kernel_id, dtype_np, local = ops
    e2 = EPS2_DEFAULT[precision.lower()] if eps2 is None else float(eps2)
    mass_t, mass_s, n_src = _limit_sources(mass_t, mass_s, N, n_sources, dtype_np, pos_t.device)
    acc, _ = _run_direct(pos_t, mass_t, mass_s, h_t, h_s, N, prec, kernel_id, MODE_FORCE, G, e2, dtype_np, group, n_src, algorithm, gather, local)
    if return_cupy:
        return acc.as_subclass(DeviceArray)
    return _to_host(acc)


def compute_nbody_potential_gpu(
    pos,
    mass,
    softening=0.0,
    G: float = G_DEFAULT,
    precision: Literal["float32", "float64", "float32_kahan"] = "float32_kahan",
    kernel: KERNEL_TYPES = "spline",
    return_cupy: bool = False,
    skip_validation: bool = False,
    *,
    eps2: float | None = None,
    device=None,
    group=None,
    n_sources=None,
    algorithm=None,
    gather="all",
):
    """Gravitational potential at each particle, ``phi_i = G sum_{j != i} m_j Phi(r_ij)``.

    Drop-in for reference ``compute_nbody_potential_gpu`` (fields.py:866-1018);
    returns (N,) in the kernel dtype.  ``dehnen_k2`` follows the reference GPU
    formula (cuda_kernels.py:394-402), which differs from the reference CPU one
    (fields.py:280-288) — see DESIGN.md.
    """
    _torch()
    ops = _prepare(pos, mass, softening, precision, kernel, skip_validation, device, group, n_sources)
    pos_t, mass_t, mass_s, h_t, h_s, N, prec, kernel_id, dtype_np, local = ops
    e2 = EPS2_DEFAULT[precision.lower()] if eps2 is None else float(eps2)
    mass_t, mass_s, n_src = _limit_sources(mass_t, mass_s, N, n_sources, dtype_np, pos_t.device)
    _, phi = _run_direct(pos_t, mass_t, mass_s, h_t, h_s, N, prec, kernel_id, MODE_POT, G, e2, dtype_np, group, n_src, algorithm, gather, local)
    if return_cupy:
        return phi.as_subclass(DeviceArray)
    return _to_host(phi)


def compute_nbody_forces_potential_gpu(
    pos,
    mass,
    softening=0.0,
    G: float = G_DEFAULT,
    precision: Literal["float32", "float64", "float32_kahan"] = "float32_kahan",
    kernel: KERNEL_TYPES = "spline",
    return_cupy: bool = False,
    skip_validation: bool = False,
    *,
    eps2: float | None = None,
    device=None,
    group=None,
    n_sources=None,
    algorithm=None,
    gather="all",
):
    """Accelerations AND potential from one fused sweep (new; the reference needs two
    O(N^2) passes, run.py:611-634).  Returns ``(acc (N,3), phi (N,))``."""
    _torch()
    ops = _prepare(pos, mass, softening, precision, kernel, skip_validation, device, group, n_sources)
    pos_t, mass_t, mass_s, h_t, h_s, N, prec, kernel_id, dtype_np, local = ops
    e2 = EPS2_DEFAULT[precision.lower()] if eps2 is None else float(eps2)
    mass_t, mass_s, n_src = _limit_sources(mass_t, mass_s, N, n_sources, dtype_np, pos_t.device)
    acc, phi = _run_direct(pos_t, mass_t, mass_s, h_t, h_s, N, prec, kernel_id, MODE_BOTH, G, e2, dtype_np, group, n_src, algorithm, gather, local)
    if return_cupy:
        return acc.as_subclass(DeviceArray), phi.as_subclass(DeviceArray)
    return _to_host(acc), _to_host(phi)


def get_gpu_info() -> dict:
    """Same keys as reference ``get_gpu_info`` (fields.py:1149-1185)."""
    try:
        torch = _torch()
    except ImportError:
        return {"available": False}
    try:
        lib = _lib.load()
        sm, maj, mnr, clk = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        tot, free = C.c_size_t(), C.c_size_t()
        name = C.create_string_buffer(256)
        with torch.cuda.device(torch.cuda.current_device()):
            torch.cuda.init()
            _lib.check(lib.nb_device_info(C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(tot), C.byref(free),
                                          C.byref(clk), name, 256))
        return {
            "available": True,
            "device_name": name.value.decode("utf-8"),
            "compute_capability": f"{maj.value}{mnr.value}",
            "memory_total": int(tot.value),
            "memory_free": int(free.value),
            "sm_count": int(sm.value),
        }
    except Exception as e:  # noqa: BLE001 - mirror the reference: report, don't raise
        print(e)
        return {"available": False, "error": str(e)}


__all__ = [
    "compute_nbody_forces_gpu",
    "compute_nbody_potential_gpu",
    "compute_nbody_forces_potential_gpu",
    "get_gpu_info",
    "G_DEFAULT",
    "KERNEL_MAP",
    "DeviceArray",
]
