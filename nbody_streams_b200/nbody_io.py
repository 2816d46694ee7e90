"""Snapshot / restart files in the reference's on-disk formats.

Only the writer side the integrator loop needs; schemas
follow reference nbody_streams/nbody_io.py so that its ``ParticleReader`` and
``continue_run`` keep working on files written here:

* ``restart.npz``  — keys ``phase_space, time, step, snapshot_counter`` and, for
  multi-species runs, ``mass_arr, softening_arr, species_names, species_N``
  (reference nbody_io.py:930-986, loader :989-1035);
* ``snapshot.h5`` / ``snapshot.NNN.h5`` — group ``snapshots/snap.###`` (gzip
  float64 (N,6)) with ``snap_time.###`` attributes, and ``properties/<species>/
  {N, m | m_array, eps | eps_array}``, ``properties/time_step``
  (reference nbody_io.py:770-928).  Needs h5py, imported lazily;
* ``snapshot.times`` — text table "snap_index time" (reference nbody_io.py:1037-1074).
"""
from __future__ import annotations

import math
import os
import queue
import threading
from pathlib import Path

import numpy as np


def _per_particle(value, n):
    if np.isscalar(value):
        return np.full(n, float(value), dtype=np.float64)
    return np.asarray(value, dtype=np.float64)


def _snapshot_filename(output_dir: Path, snap_index: int, single_file, num_files_to_write,
                       total_expected_snapshots) -> Path:
    if single_file is None:
        single_file = num_files_to_write is None
    n_files = int(num_files_to_write) if (num_files_to_write is not None and num_files_to_write > 0) else 1
    if single_file or n_files == 1:
        return output_dir / "snapshot.h5"
    if total_expected_snapshots is not None and total_expected_snapshots > 0:
        per_file = math.ceil(total_expected_snapshots / n_files)
        idx = min(int(snap_index) // per_file, n_files - 1)
    else:
        idx = int(snap_index) % n_files
    return output_dir / f"snapshot.{idx:03d}.h5"


def _save_snapshot(phase_space, snap_index, time, output_dir, *, species=None, num_dark=None, num_star=None,
                   mass_dark=None, mass_star=None, time_step=None, eps_dark=None, eps_star=None,
                   single_file=None, num_files_to_write=None, total_expected_snapshots=None) -> None:
    """Append ``snap.<index>`` to the HDF5 snapshot file (never overwrites an existing dataset)."""
    try:
        import h5py
    except ImportError as e:
        raise ImportError("h5py is required to write snapshots (save_snapshots=True)") from e

    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    fname = _snapshot_filename(output_dir, snap_index, single_file, num_files_to_write, total_expected_snapshots)
    with h5py.File(fname, "a") as f:
        snaps = f.require_group("snapshots")
        name = f"snap.{snap_index:03d}"
        if name in snaps:
            return
        snaps.create_dataset(name, data=phase_space, compression="gzip")
        snaps.attrs[f"snap_time.{snap_index:03d}"] = float(time)
        props = f.require_group("properties")
        if species is not None:
            if "n_species" not in props.attrs:
                props.attrs["n_species"] = len(species)
                props.attrs["species_names"] = np.array([s.name.encode("utf-8") for s in species])
            for s in species:
                if s.name in props:
                    continue
                grp = props.create_group(s.name)
                grp.create_dataset("N", data=int(s.N))
                for values, scalar_key, array_key in ((_per_particle(s.mass, s.N), "m", "m_array"),
                                                      (_per_particle(s.softening, s.N), "eps", "eps_array")):
                    # the reference's _is_uniform: equal to the first element within rtol 1e-10 (nbody_io.py:36-45)
                    if values.size and np.allclose(values, float(values.flat[0]), rtol=1e-10, atol=0.0):
                        grp.create_dataset(scalar_key, data=float(values.flat[0]))
                    else:
                        grp.create_dataset(array_key, data=values, compression="gzip")
        else:
            n = phase_space.shape[0]
            if num_dark is None and num_star is None:
                num_dark, num_star = n, 0
            elif num_star is None:
                num_star = n - int(num_dark)
            for key, cnt, m, eps in (("dark", num_dark, mass_dark, eps_dark), ("star", num_star, mass_star, eps_star)):
                if key not in props:
                    grp = props.create_group(key)
                    grp.create_dataset("N", data=int(cnt))
                    grp.create_dataset("m", data=float(1.0 if m is None else m))
                    grp.create_dataset("eps", data=float(0.0 if eps is None else eps))
        if "time_step" not in props:
            props.create_dataset("time_step", data=float(0.0 if time_step is None else time_step))


def _save_restart(phase_space, time, step, output_dir, snapshot_counter, *, mass_arr=None, softening_arr=None,
                  species_names=None, species_N=None) -> None:
    """Write ``restart.npz`` (compressed) for crash recovery / ``continue_run=True``."""
    out = Path(output_dir)
    out.mkdir(parents=True, exist_ok=True)
    payload = {
        "phase_space": phase_space,
        "time": np.float64(time),
        "step": np.int64(step),
        "snapshot_counter": np.int64(snapshot_counter),
    }
    if mass_arr is not None:
        payload["mass_arr"] = np.asarray(mass_arr, dtype=np.float64)
    if softening_arr is not None:
        payload["softening_arr"] = np.asarray(softening_arr, dtype=np.float64)
    if species_names is not None:
        payload["species_names"] = np.array([n.encode("utf-8") for n in species_names])
    if species_N is not None:
        payload["species_N"] = np.array(species_N, dtype=np.int64)
    # write to a temporary file and rename it into place: a crash (or interpreter exit) in the middle of the
    # write never leaves a truncated restart.npz behind
    tmp = out / "restart.tmp.npz"
    np.savez_compressed(tmp, **payload)
    os.replace(tmp, out / "restart.npz")


def _load_restart(output_dir):
    """-> (phase_space, time, step, snapshot_counter, mass_arr, softening_arr, species_names, species_N) or None."""
    path = Path(output_dir) / "restart.npz"
    if not path.exists():
        return None
    data = np.load(path, allow_pickle=False)
    files = set(data.files)
    names = None
    if "species_names" in files:
        names = [n.decode("utf-8") if isinstance(n, (bytes, np.bytes_)) else str(n) for n in data["species_names"]]
    return (
        data["phase_space"],
        float(data["time"]),
        int(data["step"]),
        int(data["snapshot_counter"]) if "snapshot_counter" in files else 0,
        data["mass_arr"] if "mass_arr" in files else None,
        data["softening_arr"] if "softening_arr" in files else None,
        names,
        [int(n) for n in data["species_N"]] if "species_N" in files else None,
    )


def _update_snapshot_times(output_dir, snap_index: int, time: float) -> None:
    """Insert / update the row for ``snap_index`` in ``snapshot.times`` (sorted, unique indices)."""
    out = Path(output_dir)
    out.mkdir(parents=True, exist_ok=True)
    path = out / "snapshot.times"
    rows = {}
    if path.exists():
        try:
            arr = np.atleast_2d(np.loadtxt(str(path), comments="#"))
            if arr.size:
                rows = {int(r[0]): float(r[1]) for r in arr}
        except Exception:  # noqa: BLE001 - unreadable file: start over, as the reference does
            rows = {}
    rows[int(snap_index)] = float(time)
    table = np.array(sorted(rows.items()), dtype=float)
    np.savetxt(str(path), table, fmt="%d %.10e", header="snap_index time", comments="# ")


class AsyncStateWriter:
    """Writes snapshots / restart files from device state without stalling the integration loop.

    `submit(kind, xv_dev, job)`: `xv_dev` is a device (N,6) float64 tensor that will not be modified any more
    (the caller passes a fresh copy); the device->host copy goes to one of two pinned buffers on a side stream
    and `job(xv_numpy)` — the actual file write — runs in a worker thread once the copy has landed, so neither
    the D2H transfer nor the (compressing, slow) file write is on the step's critical path.  Restart files
    overwrite each other, so a queued restart job that has not started when a newer one arrives is dropped.
    `flush()` blocks until everything submitted has been written (called before run_nbody_gpu returns).
    """

    def __init__(self, torch, device, n_rows: int):
        self.torch = torch
        self.device = device
        self.stream = torch.cuda.Stream(device)
        self.buffers = [torch.empty((max(n_rows, 1), 6), dtype=torch.float64, pin_memory=True) for _ in range(2)]
        self.free = queue.Queue()
        for b in self.buffers:
            self.free.put(b)
        self.jobs: queue.Queue = queue.Queue()
        self.lock = threading.Lock()
        self.latest_restart = 0
        self.error: BaseException | None = None
        self.thread = threading.Thread(target=self._worker, name="nbody-io", daemon=True)
        self.thread.start()

    def check(self) -> None:
        """Raise (once) the exception of a failed write job: called at every submit and once per step, so a full
        disk or a permission error stops the run at the next step, as the reference's synchronous writes do."""
        if self.error is not None:
            err, self.error = self.error, None
            raise err

    def submit(self, kind: str, xv_dev, job, check: bool = True) -> None:
        """`check=False`: do not raise a stored error here (multi-rank runs raise it on all ranks together)."""
        torch = self.torch
        if check:
            self.check()
        buf = self.free.get()                       # back-pressure: at most two copies in flight
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)                # xv_dev was produced on the current stream
        with torch.cuda.stream(self.stream):
            buf[: xv_dev.shape[0]].copy_(xv_dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        xv_dev.record_stream(self.stream)
        seq = 0
        if kind == "restart":
            with self.lock:
                self.latest_restart += 1
                seq = self.latest_restart
        self.jobs.put((kind, seq, buf, xv_dev.shape[0], done, job))

    def _worker(self) -> None:
        while True:
            item = self.jobs.get()
            if item is None:
                self.jobs.task_done()
                return
            kind, seq, buf, n, done, job = item
            try:
                stale = kind == "restart" and seq < self.latest_restart
                if not stale:
                    done.synchronize()
                    job(buf[:n].numpy())
            except BaseException as e:  # noqa: BLE001 - re-raised in flush()
                self.error = e
            finally:
                if kind == "restart" and seq < self.latest_restart:
                    done.synchronize()               # the copy must have finished before the buffer is reused
                self.free.put(buf)
                self.jobs.task_done()

    def flush(self) -> None:
        self.jobs.join()
        self.check()

    def close(self, raise_errors: bool = True) -> None:
        """Drain the queue, stop the worker, then (optionally) raise a stored write error."""
        self.jobs.join()
        self.jobs.put(None)
        self.thread.join(timeout=10)
        if raise_errors:
            self.check()
