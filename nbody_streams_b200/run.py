"""KDK leapfrog integrator on the B200 engine + Plummer initial conditions.

Drop-in for the GPU half of the reference's ``nbody_streams/run.py``:

* ``run_nbody_gpu``            (reference run.py:327-775)   — same signature, defaults, return
* ``_compute_accelerations_gpu`` (reference run.py:122-214) — folded into the loop below
* ``make_plummer_sphere``      (reference run.py:1225-1325) — same seeded output, vectorised

What changes underneath: state (float64 pos/vel) stays on the device in a
:class:`~nbody_streams_b200.engine.DirectEngine`; a step is 4 kernel launches
(fused kick+drift+repack, sweep, finish, kick) instead of ~15 CuPy array
expressions plus a host->device copy of the softening array every step
(reference run.py:146 note in SURVEY.md 3B).  Launched under ``torchrun`` the
same function shards the i-particles over all ranks (one process per GPU).
"""
from __future__ import annotations

import time as pytime
from pathlib import Path
from typing import Callable, Literal

import numpy as np

from . import _lib
from ._lib import MODE_BOTH, MODE_FORCE, MODE_POT
from .fields import EPS2_DEFAULT, G_DEFAULT, KERNEL_MAP, _PRECISION_MAP, DeviceArray
from .nbody_io import (AsyncStateWriter, _load_restart, _save_restart, _save_snapshot,
                       _update_snapshot_times)
from .species import Species

NBODY_UNITS = {"kpc": 1.0, "Msun": 1.0, "kpc / (km/s)": 1.0, "km/s": 1, "G": G_DEFAULT}


# ---------------------------------------------------------------------------
# initial conditions
# ---------------------------------------------------------------------------
def make_plummer_sphere(N: int, M_total: float = 10_000, a: float = 0.01, seed: int = 42069,
                        G: float = G_DEFAULT) -> tuple[np.ndarray, np.ndarray]:
    """Plummer sphere in virial equilibrium; returns ((N,6) phase space, (N,) equal masses).

    Produces the same arrays, bit for bit, as the reference generator
    (run.py:1225-1325) for the same arguments: identical draw order from
    ``default_rng(seed)`` — radii (inverse CDF), position angles, speeds by
    Aarseth-Henon-Wielen rejection, velocity angles, centre-of-mass shift.  The
    reference runs the rejection step as a Python loop of scalar draws (~13 s per
    million particles); here the (q, g) candidate pairs are drawn in blocks and
    the generator is then re-positioned exactly behind the last candidate the
    scalar loop would have consumed, which is ~100x faster and leaves the stream
    in the same place for the velocity angles.
    """
    rng = np.random.default_rng(seed)

    u = rng.uniform(0.0, 1.0, N)
    r = a / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    cos_t = rng.uniform(-1.0, 1.0, N)
    sin_t = np.sqrt(1.0 - cos_t ** 2)
    az = rng.uniform(0.0, 2.0 * np.pi, N)
    x = r * sin_t * np.cos(az)
    y = r * sin_t * np.sin(az)
    z = r * cos_t

    v_esc = np.sqrt(2.0 * G * M_total / np.sqrt(r ** 2 + a ** 2))

    # q = v / v_esc distributed as q^2 (1 - q^2)^(7/2); envelope height 0.09375
    h_max = 0.09375
    state0 = rng.bit_generator.state
    accepted = []
    n_acc = 0
    n_trials_before = 0          # candidate pairs consumed by completed blocks
    last_trial = 0               # 1-based index of the candidate that yields the N-th acceptance
    while n_acc < N:
        need = N - n_acc
        block = max(4096, int(need * 5.5) + 1024)       # acceptance rate ~ 0.2
        d = rng.random(2 * block)
        qs = d[0::2]                                     # uniform(0, 1)      -> 0 + 1 * d
        gs = 0.0 + h_max * d[1::2]                       # uniform(0, h_max)  -> 0 + h_max * d
        ok = np.flatnonzero(gs <= qs ** 2 * (1.0 - qs ** 2) ** 3.5)
        if ok.size >= need:
            ok = ok[:need]
            last_trial = n_trials_before + int(ok[-1]) + 1
        accepted.append(qs[ok])
        n_acc += ok.size
        n_trials_before += block
    q = np.concatenate(accepted) if accepted else np.empty(0)
    # put the generator where the scalar loop would have left it: 2 doubles per candidate
    rng.bit_generator.state = state0
    rng.bit_generator.advance(2 * last_trial)
    v_mag = q * v_esc

    cos_tv = rng.uniform(-1.0, 1.0, N)
    sin_tv = np.sqrt(1.0 - cos_tv ** 2)
    az_v = rng.uniform(0.0, 2.0 * np.pi, N)
    vx = v_mag * sin_tv * np.cos(az_v)
    vy = v_mag * sin_tv * np.sin(az_v)
    vz = v_mag * cos_tv

    x -= x.mean(); y -= y.mean(); z -= z.mean()
    vx -= vx.mean(); vy -= vy.mean(); vz -= vz.mean()

    phase_space = np.column_stack([x, y, z, vx, vy, vz])
    masses = np.full(N, M_total / N, dtype=np.float64)
    return phase_space, masses


# ---------------------------------------------------------------------------
# external acceleration hooks (reference run.py:197-212, :689-692)
# ---------------------------------------------------------------------------
class _PinnedStage:
    """Two page-locked (n,3) float64 buffers per direction for host potentials: the device->host copy of the
    positions and the host->device copy of the accelerations are asynchronous copies at PCIe speed, and a
    buffer is only reused two evaluations later (its previous H2D copy has been waited for by then)."""

    def __init__(self, torch, n):
        self.torch, self.k = torch, 0
        self.pos = [torch.empty((max(n, 1), 3), dtype=torch.float64, pin_memory=True) for _ in range(2)]
        self.acc = [torch.empty((max(n, 1), 3), dtype=torch.float64, pin_memory=True) for _ in range(2)]

    def next(self):
        self.k ^= 1
        return self.pos[self.k], self.acc[self.k]


def _external_acc(engine, external_potential, time, stage=None):
    """(n_local,3) float64 device tensor of the external-potential acceleration of THIS rank's particles.

    An external potential is a function of position only (reference run.py:197-212 evaluates
    ``pot.force(pos, t=)`` row by row), so each rank evaluates its own rows: no gather, no redundant CPU work.
    Runs on torch's current stream (the hook stream of run_nbody_gpu)."""
    torch = engine.torch
    n = engine.n_local
    if getattr(external_potential, "accepts_device_arrays", False):
        # device potential (PotentialGPU-like): evaluate on the device, no host round trip
        acc = external_potential.force(engine.pos.as_subclass(DeviceArray), t=time)
        if not isinstance(acc, torch.Tensor) and hasattr(acc, "__cuda_array_interface__"):
            torch.cuda.synchronize(engine.device)    # produced by another library on ITS stream (e.g. CuPy)
        return torch.as_tensor(acc, device=engine.device).to(torch.float64).contiguous()
    # host potential (agama.Potential-like): host round trip as in the reference, through pinned buffers
    cur = torch.cuda.current_stream(engine.device)
    if stage is None:
        pos_cpu = engine.pos.cpu().numpy()
        acc_np = np.asarray(external_potential.force(pos_cpu, t=time), dtype=np.float64)
        return torch.from_numpy(np.ascontiguousarray(acc_np)).to(engine.device)
    pos_pin, acc_pin = stage.next()
    pos_pin[:n].copy_(engine.pos, non_blocking=True)
    cur.synchronize()                                # only this (hook) stream: the sweep keeps running
    acc_np = np.asarray(external_potential.force(pos_pin[:n].numpy(), t=time), dtype=np.float64)
    acc_pin[:n].copy_(torch.from_numpy(np.ascontiguousarray(acc_np).reshape(n, 3)))
    return acc_pin[:n].to(engine.device, non_blocking=True)


def run_nbody_gpu(
    phase_space: np.ndarray,
    masses: np.ndarray,
    time_start: float,
    time_end: float,
    dt: float,
    softening,
    G: float = G_DEFAULT,
    precision: Literal["float32", "float64", "float32_kahan"] = "float32_kahan",
    kernel: Literal["newtonian", "plummer", "dehnen_k1", "dehnen_k2", "spline"] = "spline",
    external_potential=None,
    external_update_interval: int = 1,
    force_extra: Callable | None = None,
    output_dir: str = "./output",
    save_snapshots: bool = True,
    snapshots: int = 10,
    num_files_to_write: int = 1,
    restart_interval: int = 1000,
    continue_run: bool = False,
    overwrite: bool = False,
    verbose: bool = True,
    debug_energy: bool = False,
    species: list[Species] | None = None,
    *,
    device=None,
    group=None,
    overlap: bool = True,
    diagnostics: list | None = None,
    diagnostics_every: int = 0,
    cuda_graph: bool | None = None,
    exchange: str = "nccl",
    async_io: bool = True,
    algorithm: str | None = None,
    timeline: list | None = None,
    engine_out: list | None = None,
) -> np.ndarray:
    """GPU direct-summation N-body run with the KDK leapfrog of the reference.

    Positional / keyword arguments, defaults, snapshot schedule, restart files,
    printed lines and the returned (N,6) float64 array follow reference
    ``run_nbody_gpu`` (run.py:327-775).  KDK order per step: half kick, drift,
    ``time += dt``, forces, ``force_extra``, half kick (run.py:672-695).

    Differences (documented in DESIGN.md): masses stay float64 for
    ``precision='float64'`` (the reference rounds them to float32, run.py:502,
    which would cap agreement with its own CPU path at ~6e-8); ``debug_energy``
    takes phi from the same fused sweep as the forces instead of a second
    O(N^2) pass.

    Keyword-only extensions: ``device`` (CUDA device), ``group``
    (torch.distributed group; under torchrun the i-particles are sharded over
    its ranks), ``overlap`` (sweep the local shard while the all-gather is in
    flight; False = global summation order), ``diagnostics`` (list that
    receives ``(step, time, KE, PE, px, py, pz)`` every ``diagnostics_every``
    steps, from the fused sweep), ``cuda_graph`` (replay the whole step as one
    CUDA graph; default: on for a single rank without hooks and N <= 131072,
    where launch overhead would otherwise dominate the step), ``exchange`` (multi-GPU only:
    ``"nccl"`` all-gathers the packed j-records each step, ``"p2p"`` fuses the exchange into the
    kick+drift kernel, which stores the refreshed records into every rank's buffer over NVLink),
    ``async_io`` (snapshots / periodic restart files are copied to pinned memory on a side stream and
    written by a worker thread, so file I/O never stalls the step; False = synchronous as in the reference),
    ``algorithm`` (``'auto'``: the pair-symmetric sweep — every unordered pair evaluated once and applied to
    both particles — from ``nb_sym_min_n()`` particles on, the one-sided sweep below; ``'symmetric'`` /
    ``'onesided'`` force one; only the one-sided sweep is bitwise reproducible from run to run),
    ``timeline`` (a list that receives, per step with hooks, a dict of timing-enabled CUDA events
    ``hook_begin`` / ``hook_end`` (hook stream) and ``sweep_end`` (main stream): evidence that the hooks
    overlap the sweep), ``engine_out`` (a list that receives the DirectEngine of the run — launch counts and
    device state for benchmarks and tests; without it the engine and its device buffers are released on return).
    """
    torch = _lib.require_cuda()   # ImportError when there is no CUDA device / library (reference: "CuPy required")

    phase_space = np.asarray(phase_space, dtype=np.float64)
    if phase_space.ndim != 2 or phase_space.shape[1] != 6:
        raise ValueError(f"phase_space must be (N, 6), got {phase_space.shape}")
    prec_key = precision.lower()
    if prec_key not in _PRECISION_MAP:
        raise ValueError(f"Precision must be one of {list(_PRECISION_MAP.keys())}")
    if kernel.lower() not in KERNEL_MAP:
        raise ValueError(f"Invalid kernel '{kernel}'. Must be one of: {list(KERNEL_MAP.keys())}")
    prec, dtype_np = _PRECISION_MAP[prec_key]
    # reference run.py:502 casts to float32 unconditionally; float64 mode keeps float64 here (see docstring)
    masses = np.asarray(masses, dtype=np.float64 if prec_key == "float64" else np.float32)
    N = phase_space.shape[0]
    if masses.shape[0] != N:
        raise ValueError(f"masses must have length N={N}, got {masses.shape[0]}")
    if not np.isscalar(softening):
        softening = np.asarray(softening, dtype=dtype_np)
        if softening.shape[0] != N:
            raise ValueError(f"softening must have length N={N}, got {softening.shape[0]}")

    if save_snapshots:
        try:
            import h5py  # noqa: F401  (the reference imports it at module load, run.py:29)
        except ImportError as e:
            raise ImportError("h5py is required to write snapshots (save_snapshots=True)") from e

    from .engine import DirectEngine, gather_rows, group_info

    rank, world = group_info(group)
    is_root = rank == 0
    verbose = verbose and is_root
    output_path = Path(output_dir)

    # Root decides (it owns the output directory) and every rank learns the decision, so that a refusal raises
    # on ALL ranks instead of leaving the others blocked in the first collective; the restart payload is
    # broadcast too, so ranks with a non-shared output_dir resume from the same step.
    def _root_says(value):
        if world == 1:
            return value
        import torch.distributed as dist

        box = [value]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return box[0]

    refusal = None
    if save_snapshots and not continue_run and is_root:
        existing = sorted(output_path.glob("snapshot*.h5"))
        if existing:
            if overwrite:
                for f in existing:
                    f.unlink()
                if verbose:
                    print(f"Removed {len(existing)} existing snapshot file(s) in '{output_dir}'.")
            else:
                refusal = (f"Output directory '{output_dir}' already contains snapshot files: "
                           f"{[f.name for f in existing]}. "
                           "Pass overwrite=True to delete them, or continue_run=True to resume.")
    refusal = _root_says(refusal)
    if refusal is not None:
        raise FileExistsError(refusal)

    start_step = 0
    time = time_start
    snapshot_counter = None
    xv = phase_space
    if continue_run:
        restart_data = _root_says(_load_restart(output_path) if is_root else None)
        if restart_data is not None:
            xv, time, start_step, saved_snap_counter = restart_data[:4]
            snapshot_counter = int(saved_snap_counter)
            if verbose:
                print(f"✓ Resuming from step {start_step}, time {time:.6e}")

    _snap_kwargs: dict = dict(num_files_to_write=num_files_to_write, total_expected_snapshots=snapshots)
    _restart_kwargs: dict = {}
    if species is not None:
        _snap_kwargs["species"] = species
        _snap_kwargs["time_step"] = dt
        _soft_arr = (np.full(N, float(softening), dtype=np.float64) if np.isscalar(softening)
                     else np.asarray(softening, dtype=np.float64))
        _restart_kwargs = dict(mass_arr=masses.astype(np.float64), softening_arr=_soft_arr,
                               species_names=[s.name for s in species], species_N=[s.N for s in species])
    else:
        _snap_kwargs["mass_dark"] = float(masses[0])

    total_steps = int(round((time_end - time_start) / dt))
    remaining_steps = total_steps - start_step
    if snapshots > 1:
        snapshot_steps = np.round(np.linspace(0, total_steps, snapshots)).astype(int)
    else:
        snapshot_steps = np.array([total_steps], dtype=int)
    if snapshot_counter is None:
        snapshot_counter = int(np.searchsorted(snapshot_steps, start_step, side="left"))

    if verbose:
        print("=" * 80)
        print("GPU N-body Integration")
        print("=" * 80)
        print(f"Particles: {N:,}")
        if species is not None:
            for s in species:
                print(f"  [{s.name}] N={s.N:,}")
        print(f"Time: {time_start:.3e} -> {time_end:.3e} (dt={dt:.3e})")
        print(f"Steps: {total_steps:,} ({remaining_steps:,} remaining)")
        print(f"Kernel: {kernel}, Softening: {softening if np.isscalar(softening) else 'variable'}")
        print(f"External potential: {'Yes' if external_potential is not None else 'No'}")
        if external_potential is not None:
            print(f"  Update interval: every {external_update_interval} steps")
        print(f"Snapshots: {snapshots} (every ~{total_steps // max(snapshots, 1):,} steps)")
        print(f"Restart files: every {restart_interval} steps")
        if world > 1:
            print(f"Ranks: {world} (i-particles sharded, overlap={'on' if overlap else 'off'})")
        print("=" * 80)
        print("\nTransferring data to GPU...")

    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    eng = DirectEngine(N, masses, softening, prec, KERNEL_MAP[kernel.lower()], G, EPS2_DEFAULT[prec_key], device,
                       group=group, overlap=overlap, exchange=exchange, algorithm=algorithm)
    eng.load_state(xv)

    want_phi = bool(debug_energy or diagnostics is not None)
    mode = MODE_BOTH if want_phi else MODE_FORCE
    hooks = external_potential is not None or force_extra is not None
    cached_ext = None

    def full_pos():
        return gather_rows(eng.pos, eng.layout, eng.rank, eng.group)

    def full_vel():
        return gather_rows(eng.vel, eng.layout, eng.rank, eng.group)

    host_potential = external_potential is not None and not getattr(external_potential, "accepts_device_arrays", False)
    stage = _PinnedStage(torch, eng.n_local) if host_potential else None

    def extra_acc(step, t):
        """External-potential acceleration (cached for external_update_interval steps) of this rank's rows."""
        nonlocal cached_ext
        if external_potential is None:
            return None
        if step % external_update_interval == 0 or cached_ext is None:
            cached_ext = _external_acc(eng, external_potential, t, stage)
        return cached_ext

    def add_force_extra(total, t):
        # force_extra(pos, vel, masses, t) sees ALL particles (reference run.py:689-692): rows are gathered
        fe = force_extra(full_pos().as_subclass(DeviceArray), full_vel().as_subclass(DeviceArray), masses, t)
        if not isinstance(fe, torch.Tensor):
            if hasattr(fe, "__cuda_array_interface__"):
                torch.cuda.synchronize(eng.device)   # produced by another library on ITS stream (e.g. CuPy)
            else:
                fe = np.asarray(fe)           # host result: one H2D copy, as cp.asarray in the reference
        fe = torch.as_tensor(fe, device=eng.device).to(torch.float64)[eng.layout.rows(eng.rank)]
        return fe.contiguous() if total is None else (total + fe).contiguous()

    # Hooks run on their own stream: positions are final once the drift kernel has run, so the device->host copy
    # for a host potential, the CPU evaluation (agama-like: 10-20 ms in the reference's own docs) and the
    # host->device copy of the result all overlap the O(N^2) sweep queued on the main stream.  The reference
    # serialises them (run.py:197-212, :689-692).  Hook code must launch its device work on torch's CURRENT stream
    # (torch does by default; results of other libraries are synchronised device-wide before use).
    # With several ranks the external potential still runs on the side stream (local rows, no collective);
    # force_extra needs the gathered arrays and therefore stays on the main stream, after the sweep is queued.
    hook_stream = torch.cuda.Stream(eng.device) if hooks else None
    ev_drift = torch.cuda.Event() if hooks else None

    def eval_hooks(step, t, with_force_extra):
        cur = torch.cuda.current_stream(eng.device)
        rec = None
        if timeline is not None:
            rec = {"step": step, "hook_begin": torch.cuda.Event(enable_timing=True),
                   "hook_end": torch.cuda.Event(enable_timing=True), "sweep_end": torch.cuda.Event(enable_timing=True)}
            rec["sweep_end"].record(cur)                     # the sweep + finish are already queued on `cur`
        hook_stream.wait_event(ev_drift)
        with torch.cuda.stream(hook_stream):
            if rec is not None:
                rec["hook_begin"].record(hook_stream)
            total = extra_acc(step, t)
            if with_force_extra and world == 1:
                total = add_force_extra(total, t)
            if rec is not None:
                rec["hook_end"].record(hook_stream)
        cur.wait_stream(hook_stream)
        if total is not None:
            total.record_stream(cur)
        if with_force_extra and world > 1:
            total = add_force_extra(total, t)
        if rec is not None:
            timeline.append(rec)
        return total

    if verbose:
        print("Computing initial forces...")
    if hooks:
        ev_drift.record(torch.cuda.current_stream(eng.device))
    eng.compute(mode, positions_changed=False)
    acc_ext = eval_hooks(start_step, time, False) if hooks else None
    # the reference applies force_extra only inside the loop (run.py:689-692), not to the initial acceleration

    E_ref = 0.0

    def energies():
        e = eng.energy(with_phi=True)
        return float(e[0]), float(e[1]), e

    if want_phi:
        KE0, PE0, e = energies()
        E_ref = KE0 + PE0
        if diagnostics is not None:
            diagnostics.append((start_step, time, KE0, PE0, float(e[2]), float(e[3]), float(e[4])))
        if verbose and debug_energy:
            print(f"  [Energy t=0] KE={KE0:.4e}  PE={PE0:.4e}  E={E_ref:.4e}")

    if verbose:
        print("\nStarting integration...")

    def host_state():
        return eng.state_numpy()

    writer = AsyncStateWriter(torch, eng.device, N) if (async_io and is_root) else None

    class _WriteFailed(Exception):
        """Several ranks: the root's verdict that a file write has failed, raised on EVERY rank at the same point."""

        def __init__(self, error):
            super().__init__(str(error))
            self.error = error

    def agree_on_io(err):
        """Several ranks: only the root writes files, so only the root can see a write fail.  The root's status is
        broadcast at every write event (all ranks call write_state at the same steps) and a failure is raised on
        all of them — otherwise the root would leave the loop alone and the others would block in the next
        collective.  `err`: the root's own exception, if any (ignored on the other ranks)."""
        if is_root and err is None and writer is not None:
            try:
                writer.check()                    # a job queued at an earlier write event may have failed since
            except BaseException as e:  # noqa: BLE001
                err = e
        msg = _root_says(None if (err is None or not is_root) else f"{type(err).__name__}: {err}")
        if msg is not None:
            raise _WriteFailed(err if (is_root and err is not None)
                               else RuntimeError(f"file write failed on the root rank: {msg}"))

    def write_state(kind, counter, t, step):
        """Snapshot (kind='snapshot') or periodic restart (kind='restart') of the current state."""
        if kind == "snapshot":
            def job(xv_np, counter=counter, t=t):
                _save_snapshot(xv_np, counter, t, output_path, **_snap_kwargs)
                _update_snapshot_times(output_path, counter, t)
        else:
            def job(xv_np, t=t, step=step, counter=counter):
                _save_restart(xv_np, t, step, output_path, counter, **_restart_kwargs)
        err = None
        if async_io:
            xv_dev = eng.state_device(root_only=True)     # only the writing rank receives the rows
            if is_root:
                writer.submit(kind, xv_dev, job, check=(world == 1))
        else:
            xv_np = host_state()
            if is_root:
                try:
                    job(xv_np)
                except Exception as e:  # noqa: BLE001 - several ranks: raised on all of them by agree_on_io
                    if world == 1:
                        raise
                    err = e
        if world > 1:
            agree_on_io(err)

    t_start = pytime.perf_counter()
    _report_every = max(1, remaining_steps // 20)

    if cuda_graph is None:
        cuda_graph = world == 1 and not hooks and N <= 131072
    graph_step = eng.capture_step(mode, dt) if (cuda_graph and world == 1 and not hooks and remaining_steps > 0
                                                 and eng.n_local > 0) else None

    # Error handling.  A failed file write of the async worker is noticed at the end of the step in which it
    # failed: the loop stops there, the (consistent) state of that step is written to restart.npz synchronously,
    # and the I/O error is raised — the reference's synchronous writes fail at the same point.  If the loop itself
    # raises (a hook, Ctrl-C) the state may be in the middle of a step, so no new restart is written; the writer
    # is drained and stopped (restart.npz is replaced atomically, so the last periodic one is intact) and the
    # error propagates.
    loop_error = io_error = None
    steps_done = start_step
    try:
        if snapshot_counter < len(snapshot_steps) and snapshot_steps[snapshot_counter] == start_step:
            if save_snapshots:
                write_state("snapshot", snapshot_counter, time, start_step)
                if verbose:
                    print(f"Saved snapshot snap: {snapshot_counter:03d} at step {start_step}, time {time:.6e}...")
            snapshot_counter += 1

        for step_i in range(1, remaining_steps + 1):
            current_step = start_step + step_i

            if graph_step is not None:
                graph_step()                     # the same four launches, replayed as one CUDA graph
                time += dt
            else:
                eng.kick_drift(dt, acc_ext)          # half kick + drift (+ repack own shard), fused
                if hooks:
                    ev_drift.record(torch.cuda.current_stream(eng.device))
                time += dt
                eng.compute(mode)                    # all-gather (world>1) + sweep + finish (asynchronous)
                if hooks:                            # evaluated while the sweep runs
                    acc_ext = eval_hooks(current_step, time, force_extra is not None)
                eng.kick(dt, acc_ext)                # half kick
            steps_done = current_step                # the state is that of this step from here on

            while snapshot_counter < len(snapshot_steps) and current_step >= snapshot_steps[snapshot_counter]:
                if save_snapshots:
                    write_state("snapshot", snapshot_counter, time, current_step)
                    if verbose:
                        print(f"Saved snapshot id={snapshot_counter:03d} at step {current_step}, time {time:.6e}...",
                              flush=True)
                snapshot_counter += 1

            if diagnostics is not None and diagnostics_every and (step_i % diagnostics_every == 0 or
                                                                  step_i == remaining_steps):
                KE, PE, e = energies()
                diagnostics.append((current_step, time, KE, PE, float(e[2]), float(e[3]), float(e[4])))

            if (verbose or debug_energy) and step_i % _report_every == 0:
                line = None
                if verbose:
                    elapsed = pytime.perf_counter() - t_start
                    rate = step_i / elapsed if elapsed > 0 else 0
                    eta = (remaining_steps - step_i) / rate if rate > 0 else 0
                    line = (f"  Step {current_step:>6}/{total_steps} | t={time:.4e} | "
                            f"Snapshots: {snapshot_counter}/{len(snapshot_steps)} | {rate:.1f} steps/s | "
                            f"avg {elapsed / step_i * 1000:.1f}ms/step | ETA {eta:.0f}s")
                if debug_energy and E_ref != 0.0:
                    KE, PE, _ = energies()          # collective: every rank takes part
                    dE = (KE + PE - E_ref) / abs(E_ref)
                    Q = KE / abs(PE) if PE != 0.0 else float("nan")
                    if line is not None:
                        line += f" | Q={Q:.3f}  dE/E={dE:+.2e}"
                if line is not None:
                    print(line, flush=True)

            if current_step > 0 and (current_step % restart_interval) == 0:
                write_state("restart", snapshot_counter, time, current_step)
            if writer is not None and world == 1:
                try:
                    writer.check()
                except BaseException as e:  # noqa: BLE001 - raised after the restart file is safe
                    io_error = e
                    break

        if io_error is None and snapshot_counter < len(snapshot_steps) and snapshot_steps[-1] == total_steps:
            if save_snapshots:
                if verbose:
                    print(f"<=Saving final snapshot snap.{snapshot_counter:03d} at step {total_steps}, "
                          f"time {time:.6e}=>")
                write_state("snapshot", snapshot_counter, time, total_steps)
            snapshot_counter += 1

    except _WriteFailed as e:           # several ranks: every rank arrives here at the same step
        io_error = e.error
    except BaseException as e:  # noqa: BLE001 - re-raised below, once the writer thread is stopped
        loop_error = e
    if loop_error is not None:
        if writer is not None:
            writer.close(raise_errors=False)
        raise loop_error
    xv_final_cpu = host_state()
    if writer is not None:
        try:
            writer.close()                      # everything queued is on disk before the final restart is written
        except BaseException as e:  # noqa: BLE001
            io_error = io_error or e
    if is_root:
        try:
            _save_restart(xv_final_cpu, time, steps_done, output_path, snapshot_counter, **_restart_kwargs)
        except Exception as e:  # noqa: BLE001
            if world == 1 or io_error is not None:
                raise
            io_error = e
    if world > 1:
        # the outcome of the run is the root's: a write that failed at the very end is raised on every rank too
        msg = _root_says(None if (io_error is None or not is_root) else f"{type(io_error).__name__}: {io_error}")
        if msg is not None and io_error is None:
            io_error = RuntimeError(f"file write failed on the root rank: {msg}")
    if io_error is not None:
        raise io_error

    if verbose:
        total_time = pytime.perf_counter() - t_start
        avg = total_time / remaining_steps if remaining_steps > 0 else 0
        print("\n" + "=" * 80)
        print("Integration Complete")
        print("=" * 80)
        print(f"Final time: {time:.6e}")
        print(f"Total wall time: {total_time:.2f} s")
        print(f"Steps per second: {remaining_steps / total_time:.1f}" if total_time > 0 else "Steps per second: n/a")
        print(f"Average step time: {avg * 1000:.1f} ms")
        print(f"Snapshots saved: {snapshot_counter}")
        print("=" * 80)

    if engine_out is not None:
        engine_out.append(eng)
    return xv_final_cpu


__all__ = ["run_nbody_gpu", "make_plummer_sphere", "G_DEFAULT", "NBODY_UNITS"]
