"""GPU integrator tests: run_nbody_gpu / run_simulation against the oracle's KDK loop and the golden
runs of the reference (run_simulation(architecture='cpu', method='direct'))."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nb():
    import nbody_streams_b200 as nb

    return nb


def test_kdk_float64_matches_oracle_loop(nb, tmp_path):
    xv, m = nb.make_plummer_sphere(2000, M_total=1e9, a=1.0, seed=11)
    final = nb.run_nbody_gpu(xv, m, 0.0, 30e-4, 1e-4, 0.05, precision="float64", kernel="spline",
                             output_dir=str(tmp_path), save_snapshots=False, verbose=False)
    ref, _ = oracle.run_nbody_cpu_direct(xv, m, 0.0, 30e-4, 1e-4, 0.05, kernel="spline")
    assert final.shape == (2000, 6) and final.dtype == np.float64
    assert np.abs(final - ref).max() / np.abs(ref).max() < 1e-12
    # the final restart file is always written (reference run.py:754-757)
    r = np.load(tmp_path / "restart.npz")
    assert np.array_equal(r["phase_space"], final) and int(r["step"]) == 30


def test_c1_config_100_steps_vs_reference_golden(nb, golden, tmp_path):
    """BASELINE config 1 (Plummer N=10k, spline h=0.05, float64, 100 KDK steps dt=1e-4): final state and
    energy drift of the GPU run against the reference's own CPU run."""
    g = golden("c1_plummer10k.npz")
    xv, m = nb.make_plummer_sphere(10_000, M_total=1e9, a=1.0, seed=42069)
    diags = []
    res = nb.run_simulation(xv, [nb.Species.dark(10_000, mass=1e5, softening=0.05)], 0.0, 100 * 1e-4, 1e-4,
                            architecture="gpu", method="direct", precision="float64", output_dir=str(tmp_path),
                            save_snapshots=False, verbose=False, diagnostics=diags, diagnostics_every=50)
    final = res["dark"]
    ref_rows = g["final_rows"]
    err = np.abs(final[::10] - ref_rows).max(axis=1) / np.abs(ref_rows).max()
    assert np.quantile(err, 0.99) < 1e-11 and err.max() < 1e-8
    E0 = float(g["ke0"] + g["pe0"])
    dE_ref = float(g["ke1"] + g["pe1"]) - E0
    step0, _, ke0, pe0, *_ = diags[0]
    stepN, _, ke1, pe1, px, py, pz = diags[-1]
    assert step0 == 0 and stepN == 100
    assert abs(ke0 - float(g["ke0"])) / abs(E0) < 1e-12 and abs(pe0 - float(g["pe0"])) / abs(E0) < 1e-12
    dE = (ke1 + pe1) - (ke0 + pe0)
    assert abs(dE - dE_ref) / abs(E0) < 1e-10          # energy drift equal to the reference's (+9.9e-6)
    assert abs(dE / E0) < 1e-4
    p_scale = 1e9 * np.sqrt(2 * ke0 / 1e9)
    assert np.linalg.norm([px, py, pz]) / p_scale < 1e-12   # momentum conservation


@pytest.mark.parametrize("precision", ["float32", "float32_kahan"])
def test_float32_run_tracks_float64_reference(nb, golden, precision, tmp_path):
    # reference tests/test_physics.py:644-715: GPU float32_kahan trajectory vs CPU, rtol 1e-3 atol 1e-6
    g = golden("c1_plummer10k.npz")
    xv, m = nb.make_plummer_sphere(10_000, M_total=1e9, a=1.0, seed=42069)
    final = nb.run_nbody_gpu(xv, m, 0.0, 100 * 1e-4, 1e-4, 0.05, precision=precision, kernel="spline",
                             output_dir=str(tmp_path), save_snapshots=False, verbose=False)
    ref = g["final_rows"]
    assert np.allclose(final[::10], ref, rtol=1e-3, atol=1e-6 * np.abs(ref).max())


def test_multispecies_run_simulation_vs_reference_golden(nb, golden, tmp_path):
    g = golden("multispecies_n1251.npz")
    sp = [nb.Species.dark(1000, mass=float(g["mass"][0]), softening=0.05),
          nb.Species.star(250, mass=float(g["mass"][1000]), softening=np.linspace(0.01, 0.03, 250)),
          nb.Species("bh", 1, 1e8, softening=0.001)]
    res = nb.run_simulation(g["xv"], sp, 0.0, 20 * 1e-4, 1e-4, precision="float64", output_dir=str(tmp_path),
                            save_snapshots=False, verbose=False)
    assert list(res) == ["dark", "star", "bh"]
    assert res["dark"].shape == (1000, 6) and res["star"].shape == (250, 6) and res["bh"].shape == (1, 6)
    final = np.vstack([res["dark"], res["star"], res["bh"]])
    ref = np.vstack([g["final_dark"], g["final_star"], g["final_bh"]])
    d = np.abs(final - ref).max(axis=1) / np.abs(ref).max()
    assert np.quantile(d, 0.99) < 1e-12 and d.max() < 1e-7   # one star orbits 0.02 kpc from the BH (chaotic)


def test_energy_conservation_small_system(nb, tmp_path):
    # reference tests/test_physics.py:51-62,226-244: N=80, EPS=0.2, G=1, 50 steps: |dE/E| < 1 %, momentum 1e-8
    rng = np.random.default_rng(42)
    N = 80
    pos = rng.standard_normal((N, 3))
    vel = rng.standard_normal((N, 3)) * 0.1
    mass = np.full(N, 1.0 / N)
    vel -= (mass[:, None] * vel).sum(0) / mass.sum()
    xv = np.hstack([pos, vel])
    for precision, etol in (("float64", 1e-2), ("float32_kahan", 5e-2)):
        diags = []
        nb.run_nbody_gpu(xv, mass, 0.0, 0.25, 5e-3, 0.2, G=1.0, precision=precision, kernel="spline",
                         output_dir=str(tmp_path / precision), save_snapshots=False, verbose=False,
                         diagnostics=diags, diagnostics_every=10)
        E = np.array([d[2] + d[3] for d in diags])
        assert np.abs(E - E[0]).max() / abs(E[0]) < etol
        p = np.array([d[4:7] for d in diags])
        assert np.abs(p).max() < 1e-8 if precision == "float64" else np.abs(p).max() < 1e-6


def test_hooks_external_potential_and_force_extra(nb, tmp_path):
    class Harmonic:   # host potential with the agama call shape: force(pos, t=) -> (N,3)
        def force(self, pos, t=0.0):
            return -0.5 * np.asarray(pos)

    calls = []

    def drag(pos, vel, masses, time):
        calls.append(time)
        return -0.01 * vel.get()          # reference closures call .get() on the device arrays

    xv, m = nb.make_plummer_sphere(500, M_total=1e6, a=1.0, seed=3)
    final = nb.run_nbody_gpu(xv, m, 0.0, 10e-3, 1e-3, 0.05, precision="float64", kernel="plummer",
                             external_potential=Harmonic(), external_update_interval=1, force_extra=drag,
                             output_dir=str(tmp_path), save_snapshots=False, verbose=False)
    assert len(calls) == 10
    # host restatement of the same loop (run.py:666-695)
    x, v = xv[:, :3].copy(), xv[:, 3:].copy()
    pot = Harmonic()
    acc = oracle.compute_nbody_forces_cpu(x, m, 0.05, kernel="plummer") + pot.force(x)
    t = 0.0
    for _ in range(10):
        v += acc * 0.5e-3
        x += v * 1e-3
        t += 1e-3
        acc = oracle.compute_nbody_forces_cpu(x, m, 0.05, kernel="plummer") + pot.force(x) - 0.01 * v
        v += acc * 0.5e-3
    assert np.abs(final - np.hstack([x, v])).max() / np.abs(xv).max() < 1e-12


def test_restart_continue_run(nb, tmp_path):
    xv, m = nb.make_plummer_sphere(600, M_total=1e9, a=1.0, seed=8)
    kw = dict(precision="float64", kernel="spline", save_snapshots=False, verbose=False)
    full = nb.run_nbody_gpu(xv, m, 0.0, 20e-4, 1e-4, 0.05, output_dir=str(tmp_path / "a"), **kw)
    nb.run_nbody_gpu(xv, m, 0.0, 10e-4, 1e-4, 0.05, output_dir=str(tmp_path / "b"), **kw)
    # resume the second run to t = 20e-4 from its restart file; total_steps is recomputed from (t1 - t0)/dt
    resumed = nb.run_nbody_gpu(xv, m, 0.0, 20e-4, 1e-4, 0.05, output_dir=str(tmp_path / "b"), continue_run=True,
                               **kw)
    assert np.abs(resumed - full).max() / np.abs(full).max() < 1e-13


def test_run_error_paths(nb, tmp_path):
    xv, m = nb.make_plummer_sphere(64, seed=1)
    with pytest.raises(ValueError):
        nb.run_nbody_gpu(xv[:, :5], m, 0, 1e-3, 1e-4, 0.01, output_dir=str(tmp_path))
    with pytest.raises(ValueError):
        nb.run_nbody_gpu(xv, m[:10], 0, 1e-3, 1e-4, 0.01, output_dir=str(tmp_path))
    with pytest.raises(ValueError):
        nb.run_nbody_gpu(xv, m, 0, 1e-3, 1e-4, 0.01, precision="half", output_dir=str(tmp_path))
    sp = [nb.Species.dark(64, 1.0, 0.01)]
    with pytest.raises(TypeError):
        nb.run_simulation(xv, sp, 0, 1e-3, 1e-4, output_dir=str(tmp_path), save_snapshots=False, bogus_kw=1)
    with pytest.raises(ValueError):
        nb.run_simulation(xv, [nb.Species.dark(63, 1.0)], 0, 1e-3, 1e-4, output_dir=str(tmp_path))
    with pytest.raises(NotImplementedError):
        nb.run_simulation(xv, sp, 0, 1e-3, 1e-4, architecture="cpu", output_dir=str(tmp_path))
    (tmp_path / "snapshot.h5").write_bytes(b"x")
    try:
        import h5py  # noqa: F401
        expected = FileExistsError
    except ImportError:
        expected = ImportError          # no h5py in this image: refused up front, before any compute
    with pytest.raises(expected):
        nb.run_nbody_gpu(xv, m, 0, 1e-3, 1e-4, 0.01, output_dir=str(tmp_path), save_snapshots=True, verbose=False)


def test_cuda_graph_step_is_bitwise_identical(nb, tmp_path):
    xv, m = nb.make_plummer_sphere(3000, M_total=1e9, a=1.0, seed=21)
    kw = dict(kernel="spline", save_snapshots=False, verbose=False)
    for precision in ("float32_kahan", "float64"):
        a = nb.run_nbody_gpu(xv, m, 0.0, 25e-4, 1e-4, 0.05, precision=precision, cuda_graph=True,
                             output_dir=str(tmp_path / f"g{precision}"), **kw)
        b = nb.run_nbody_gpu(xv, m, 0.0, 25e-4, 1e-4, 0.05, precision=precision, cuda_graph=False,
                             output_dir=str(tmp_path / f"n{precision}"), **kw)
        assert np.array_equal(a, b)
    # with diagnostics (fused potential inside the graph)
    d1, d2 = [], []
    nb.run_nbody_gpu(xv, m, 0.0, 10e-4, 1e-4, 0.05, cuda_graph=True, diagnostics=d1, diagnostics_every=5,
                     output_dir=str(tmp_path / "d1"), **kw)
    nb.run_nbody_gpu(xv, m, 0.0, 10e-4, 1e-4, 0.05, cuda_graph=False, diagnostics=d2, diagnostics_every=5,
                     output_dir=str(tmp_path / "d2"), **kw)
    assert d1 == d2 and len(d1) == 3


def test_debug_energy_prints_virial_and_drift(nb, tmp_path, capsys):
    # reference tests/test_multi_species.py:714-760: debug_energy=True prints "[Energy t=0]" and "Q=... dE/E=..."
    xv, m = nb.make_plummer_sphere(400, M_total=1e8, a=1.0, seed=2)
    nb.run_nbody_gpu(xv, m, 0.0, 40e-4, 1e-4, 0.05, output_dir=str(tmp_path), save_snapshots=False, verbose=True,
                     debug_energy=True)
    out = capsys.readouterr().out
    assert "GPU N-body Integration" in out and "[Energy t=0]" in out
    assert "Q=" in out and "dE/E=" in out and "Integration Complete" in out


def test_old_api_without_species_and_scalar_softening(nb, tmp_path):
    # reference tests/test_multi_species.py:659-707: run_nbody_gpu(phase_space, masses, t0, t1, dt, softening)
    xv, m = nb.make_plummer_sphere(300, M_total=1e8, a=1.0, seed=4)
    final = nb.run_nbody_gpu(xv, m, 0.0, 5e-4, 1e-4, 0.05, output_dir=str(tmp_path), save_snapshots=False,
                             verbose=False)
    assert final.shape == (300, 6) and np.isfinite(final).all()
    r = np.load(tmp_path / "restart.npz")
    assert "mass_arr" not in r.files            # legacy restart layout without species keys
    # per-particle softening array and float32 masses are accepted as in the reference
    final2 = nb.run_nbody_gpu(xv, m.astype(np.float32), 0.0, 5e-4, 1e-4, np.full(300, 0.05), precision="float32",
                              output_dir=str(tmp_path / "b"), save_snapshots=False, verbose=False)
    assert np.allclose(final2, final, rtol=1e-4, atol=1e-6 * np.abs(final).max())


def test_run_simulation_species_counts(nb, tmp_path):
    # reference tests/test_multi_species.py:501-604: 1, 2 and 3 species incl. a single heavy particle
    xv1, m1 = nb.make_plummer_sphere(200, M_total=1e8, a=1.0, seed=1)
    xv2, m2 = nb.make_plummer_sphere(100, M_total=1e6, a=0.3, seed=2)
    cases = [
        (xv1, [nb.Species.dark(200, float(m1[0]), 0.05)]),
        (np.vstack([xv1, xv2]), [nb.Species.dark(200, float(m1[0]), 0.05), nb.Species.star(100, m2, 0.02)]),
        (np.vstack([xv1, xv2, np.zeros((1, 6))]),
         [nb.Species.dark(200, float(m1[0]), 0.05), nb.Species.star(100, float(m2[0]), np.full(100, 0.02)),
          nb.Species("bh", 1, 1e6, 0.01)]),
    ]
    for i, (xv, sp) in enumerate(cases):
        res = nb.run_simulation(xv, sp, 0.0, 3e-4, 1e-4, output_dir=str(tmp_path / str(i)), save_snapshots=False,
                                verbose=False)
        assert list(res) == [s.name for s in sp]
        for s in sp:
            assert res[s.name].shape == (s.N, 6) and res[s.name].dtype == np.float64
        r = np.load(tmp_path / str(i) / "restart.npz")
        assert [n.decode() for n in r["species_names"]] == [s.name for s in sp]
        assert list(r["species_N"]) == [s.N for s in sp] and int(r["step"]) == 3


def test_snapshot_files_when_h5py_is_available(nb, tmp_path):
    h5py = pytest.importorskip("h5py")
    xv, m = nb.make_plummer_sphere(128, M_total=1e8, a=1.0, seed=9)
    final = nb.run_simulation(xv, [nb.Species.dark(128, float(m[0]), 0.05)], 0.0, 10e-4, 1e-4,
                              output_dir=str(tmp_path), save_snapshots=True, snapshots=3, verbose=False)["dark"]
    with h5py.File(tmp_path / "snapshot.h5") as f:
        assert sorted(f["snapshots"]) == ["snap.000", "snap.001", "snap.002"]
        assert np.array_equal(f["snapshots/snap.002"][...], final)      # final array == last snapshot
        assert int(f["properties/dark/N"][()]) == 128
    t = np.loadtxt(tmp_path / "snapshot.times", comments="#")
    assert np.allclose(t[:, 1], [0.0, 5e-4, 10e-4])


def test_host_potential_overlaps_the_sweep(nb, tmp_path):
    """The host-side external potential is evaluated WHILE the O(N^2) sweep runs (hooks on a side stream;
    reference run.py:197-212 serialises them).  Asserted on the CUDA event timeline of every step, not on wall
    clock: the hook's device->host copy starts, and the host->device copy of its result ends, before the sweep
    of the same step has finished."""
    import time

    class SlowHarmonic:
        def force(self, pos, t=0.0):
            time.sleep(0.004)                  # stands for agama's CPU evaluation
            return -0.5 * np.asarray(pos)

    N, steps = 300_000, 6                       # sweep ~24 ms per step >> hook 4 ms + copies (margin for a loaded host)
    xv, m = nb.make_plummer_sphere(N, M_total=1e9, a=1.0, seed=3)
    tl = []
    out = nb.run_nbody_gpu(xv, m, 0.0, steps * 1e-4, 1e-4, 0.05, precision="float32", kernel="plummer",
                           save_snapshots=False, verbose=False, output_dir=str(tmp_path),
                           external_potential=SlowHarmonic(), timeline=tl)
    assert np.isfinite(out).all()
    assert len(tl) == steps + 1                 # initial evaluation + one per step
    for rec in tl[1:]:
        begin_to_sweep_end = rec["hook_begin"].elapsed_time(rec["sweep_end"])
        end_to_sweep_end = rec["hook_end"].elapsed_time(rec["sweep_end"])
        hook_ms = rec["hook_begin"].elapsed_time(rec["hook_end"])
        assert hook_ms >= 4.0                                 # the CPU evaluation really sat between the two events
        assert begin_to_sweep_end > hook_ms                   # the hook began at least its own length before ...
        assert end_to_sweep_end > 0.0, (rec["step"], end_to_sweep_end)   # ... and was done before the sweep ended


def test_failed_async_write_stops_the_run_and_keeps_a_restart(nb, tmp_path, monkeypatch):
    """A failing periodic restart write (disk full, permissions) is raised at the end of the step in which it
    failed — not at the end of the run — and restart.npz then holds the state of that step."""
    import nbody_streams_b200.run as run_mod

    real = run_mod._save_restart
    calls = {"n": 0}

    def flaky(xv_np, t, step, *a, **k):
        calls["n"] += 1
        if calls["n"] == 1:
            raise OSError("No space left on device (simulated)")
        return real(xv_np, t, step, *a, **k)

    monkeypatch.setattr(run_mod, "_save_restart", flaky)
    xv, m = nb.make_plummer_sphere(2000, M_total=1e9, a=1.0, seed=5)
    # a step takes ~40 us here and the failing job runs in the writer thread, so the run must be long enough for the
    # failure to be noticed in the middle of it: 20000 steps, first restart write (the failing one) at step 50
    steps, dt = 20_000, 1e-6
    kw = dict(precision="float64", kernel="spline", save_snapshots=False, verbose=False, restart_interval=50)
    with pytest.raises(OSError, match="simulated"):
        nb.run_nbody_gpu(xv, m, 0.0, steps * dt, dt, 0.05, output_dir=str(tmp_path), **kw)
    r = np.load(tmp_path / "restart.npz")
    stopped = int(r["step"])
    assert 50 <= stopped < steps                # stopped long before the end, at a step boundary
    assert not (tmp_path / "restart.tmp.npz").exists()
    # the state in the file is the state of that step: continuing from it reproduces the uninterrupted run
    monkeypatch.setattr(run_mod, "_save_restart", real)
    short = stopped + 100
    full = nb.run_nbody_gpu(xv, m, 0.0, short * dt, dt, 0.05, output_dir=str(tmp_path / "full"), **kw)
    resumed = nb.run_nbody_gpu(xv, m, 0.0, short * dt, dt, 0.05, output_dir=str(tmp_path), continue_run=True, **kw)
    assert np.abs(resumed - full).max() / np.abs(full).max() < 1e-13


def test_exception_in_a_hook_leaves_the_last_restart_intact(nb, tmp_path):
    """If the loop itself raises (here: force_extra), the writer thread is drained and stopped and the last
    periodic restart.npz is complete and loadable (it is replaced atomically)."""
    def bomb(pos, vel, masses, t):
        if t > 9.5e-4:
            raise RuntimeError("hook failed")
        return 0.0 * vel.get()

    xv, m = nb.make_plummer_sphere(1500, M_total=1e9, a=1.0, seed=6)
    with pytest.raises(RuntimeError, match="hook failed"):
        nb.run_nbody_gpu(xv, m, 0.0, 30e-4, 1e-4, 0.05, precision="float64", kernel="plummer", force_extra=bomb,
                         save_snapshots=False, verbose=False, restart_interval=3, output_dir=str(tmp_path))
    r = np.load(tmp_path / "restart.npz")
    assert int(r["step"]) in (6, 9) and r["phase_space"].shape == (1500, 6) and np.isfinite(r["phase_space"]).all()


def test_async_io_restart_files_match_synchronous(nb, tmp_path):
    """Periodic restart files written by the background writer equal the synchronous ones; the last periodic
    restart is superseded by the final one exactly as in the reference (run.py:734-757)."""
    xv, m = nb.make_plummer_sphere(5000, M_total=1e9, a=1.0, seed=13)
    kw = dict(precision="float64", kernel="spline", save_snapshots=False, verbose=False, restart_interval=3)
    a = nb.run_nbody_gpu(xv, m, 0.0, 10e-4, 1e-4, 0.05, output_dir=str(tmp_path / "async"), async_io=True, **kw)
    b = nb.run_nbody_gpu(xv, m, 0.0, 10e-4, 1e-4, 0.05, output_dir=str(tmp_path / "sync"), async_io=False, **kw)
    assert np.array_equal(a, b)
    ra, rb = np.load(tmp_path / "async" / "restart.npz"), np.load(tmp_path / "sync" / "restart.npz")
    assert int(ra["step"]) == int(rb["step"]) == 10
    assert np.array_equal(ra["phase_space"], rb["phase_space"]) and np.array_equal(ra["phase_space"], a)
    # stop after 9 steps' worth of periodic restarts: resume from the step-9 file written asynchronously
    nb.run_nbody_gpu(xv, m, 0.0, 9e-4, 1e-4, 0.05, output_dir=str(tmp_path / "part"), async_io=True, **kw)
    c = nb.run_nbody_gpu(xv, m, 0.0, 10e-4, 1e-4, 0.05, output_dir=str(tmp_path / "part"), async_io=True,
                         continue_run=True, **kw)
    assert np.abs(c - a).max() / np.abs(a).max() < 1e-13


# ---- snapshots (HDF5 schema through tests/fake_h5py.py: h5py is not installed in this image) ----------------

@pytest.fixture
def h5(monkeypatch):
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import fake_h5py

    monkeypatch.setitem(sys.modules, "h5py", fake_h5py)
    return fake_h5py


@pytest.mark.parametrize("async_io", [True, False])
def test_snapshots_follow_the_reference_schedule(nb, tmp_path, h5, async_io):
    """snapshots=5 over 20 steps -> np.round(np.linspace(0, 20, 5)) = steps 0, 5, 10, 15, 20 (run.py:571-574);
    snap.000 is the initial state, the last one the returned final state; times in the attrs and snapshot.times."""
    xv, m = nb.make_plummer_sphere(3001, M_total=1e9, a=1.0, seed=17)
    out = tmp_path / "run"
    kw = dict(precision="float64", kernel="spline", verbose=False, async_io=async_io)
    final = nb.run_nbody_gpu(xv, m, 0.0, 20e-4, 1e-4, 0.05, output_dir=str(out), save_snapshots=True, snapshots=5, **kw)
    flat = h5.flatten(out / "snapshot.h5")
    names = sorted(k for k in flat if k.startswith("/snapshots/snap.") and "#" not in k)
    assert names == [f"/snapshots/snap.{i:03d}" for i in range(5)]
    assert np.array_equal(flat["/snapshots/snap.000"], xv)
    assert np.array_equal(flat["/snapshots/snap.004"], final)
    for i, step in enumerate((0, 5, 10, 15, 20)):
        assert float(flat[f"/snapshots@snap_time.{i:03d}"]) == pytest.approx(step * 1e-4, rel=1e-12, abs=1e-18)
    assert int(flat["/properties/dark/N"]) == 3001 and float(flat["/properties/dark/m"]) == float(m[0])
    times = np.loadtxt(out / "snapshot.times")
    assert times.shape == (5, 2) and np.allclose(times[:, 1], np.arange(0, 21, 5) * 1e-4)
    # the intermediate snapshot equals a run that stops there
    mid = nb.run_nbody_gpu(xv, m, 0.0, 10e-4, 1e-4, 0.05, output_dir=str(tmp_path / "mid"), save_snapshots=False, **kw)
    assert np.array_equal(flat["/snapshots/snap.002"], mid)
    # existing snapshots: refused unless overwrite / continue_run (run.py:513-526)
    with pytest.raises(FileExistsError):
        nb.run_nbody_gpu(xv, m, 0.0, 2e-4, 1e-4, 0.05, output_dir=str(out), save_snapshots=True, **kw)
    nb.run_nbody_gpu(xv, m, 0.0, 2e-4, 1e-4, 0.05, output_dir=str(out), save_snapshots=True, snapshots=2, overwrite=True,
                     **kw)
    flat2 = h5.flatten(out / "snapshot.h5")
    assert sorted(k for k in flat2 if k.startswith("/snapshots/snap.") and "#" not in k) == \
        ["/snapshots/snap.000", "/snapshots/snap.001"]


def test_multi_species_snapshots_split_over_files_and_resume(nb, tmp_path, h5):
    xd, md = nb.make_plummer_sphere(1500, M_total=1e9, a=1.0, seed=1)
    xs, ms = nb.make_plummer_sphere(500, M_total=2e7, a=0.3, seed=2)
    xv = np.vstack([xd, xs])
    sp = [nb.Species.dark(1500, mass=float(md[0]), softening=0.05),
          nb.Species.star(500, mass=float(ms[0]), softening=np.linspace(0.01, 0.03, 500))]
    kw = dict(architecture="gpu", method="direct", verbose=False, snapshots=4, num_files_to_write=2)
    full = nb.run_simulation(xv, sp, 0.0, 12e-4, 1e-4, output_dir=str(tmp_path / "full"), **kw)
    f0, f1 = h5.flatten(tmp_path / "full" / "snapshot.000.h5"), h5.flatten(tmp_path / "full" / "snapshot.001.h5")
    assert "/snapshots/snap.000" in f0 and "/snapshots/snap.001" in f0 and "/snapshots/snap.002" in f1 \
        and "/snapshots/snap.003" in f1
    assert [n.decode() for n in f0["/properties@species_names"]] == ["dark", "star"]
    assert f0["/properties/star/eps_array"].shape == (500,) and float(f0["/properties/dark/eps"]) == 0.05
    assert np.array_equal(f1["/snapshots/snap.003"], np.vstack([full["dark"], full["star"]]))
    # stop after 8 of 12 steps, then continue_run: same snapshots, same final state
    part = tmp_path / "part"
    nb.run_simulation(xv, sp, 0.0, 8e-4, 1e-4, output_dir=str(part), **dict(kw, snapshots=3))   # steps 0, 4, 8
    res = nb.run_simulation(xv, sp, 0.0, 12e-4, 1e-4, output_dir=str(part), continue_run=True, **kw)
    assert np.array_equal(res["dark"], full["dark"]) and np.array_equal(res["star"], full["star"])
