"""Host-side logic that needs no GPU: Species helpers, input validation, the seeded Plummer generator
(bit-identical to the reference's), restart / snapshot-times files, the shard layout, the C-ABI surface."""
import ctypes as C
import hashlib
import re
import warnings
from pathlib import Path

import numpy as np
import pytest

import nbody_streams_b200 as nb
from nbody_streams_b200 import _lib, fields, nbody_io, species
from nbody_streams_b200.engine import ShardLayout

ROOT = Path(__file__).resolve().parents[1]


# ---------------------------------------------------------------- Species (reference tests/test_multi_species.py)
def test_species_validation():
    s = nb.Species.dark(N=10, mass=1e6, softening=0.1)
    assert (s.name, s.N, s.mass, s.softening) == ("dark", 10, 1e6, 0.1)
    assert nb.Species.star(5, 1.0).name == "star" and nb.Species.star(5, 1.0).softening == 0.0
    with pytest.raises(ValueError):
        nb.Species("", 1, 1.0)
    with pytest.raises(ValueError):
        nb.Species("x", 0, 1.0)
    with pytest.raises(ValueError):
        nb.Species("x", 3, np.ones(4))
    with pytest.raises(ValueError):
        nb.Species("x", 3, 1.0, softening=np.ones(2))
    nb.Species("x", 3, np.ones(3), softening=np.ones(3))


def test_build_split_validate():
    sp = [nb.Species.dark(3, 2.0, 0.1), nb.Species.star(2, np.array([1.0, 3.0]), np.array([0.01, 0.02])),
          nb.Species("bh", 1, 1e8, 0.001)]
    m, h = species._build_particle_arrays(sp)
    assert m.dtype == np.float64 and h.dtype == np.float64
    assert np.array_equal(m, [2, 2, 2, 1, 3, 1e8]) and np.array_equal(h, [0.1, 0.1, 0.1, 0.01, 0.02, 0.001])
    xv = np.arange(36.0).reshape(6, 6)
    parts = species._split_by_species(xv, sp)
    assert list(parts) == ["dark", "star", "bh"]
    assert np.shares_memory(parts["star"], xv) and np.array_equal(parts["star"], xv[3:5]) and parts["bh"].shape == (1, 6)
    species._validate_species(xv, sp)
    with pytest.raises(ValueError):
        species._validate_species(xv, [])
    with pytest.raises(ValueError):
        species._validate_species(xv, [nb.Species.dark(3, 1.0), nb.Species.dark(3, 1.0)])
    with pytest.raises(ValueError):
        species._validate_species(xv, [nb.Species.dark(5, 1.0)])


def test_performance_warnings():
    for n, arch, method, expect in [(600_000, "gpu", "direct", True), (400_000, "gpu", "direct", False),
                                    (30_000, "cpu", "direct", True), (2_500_000, "gpu", "direct", True),
                                    (2_500_000, "gpu", "tree", False)]:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            species._emit_performance_warnings(n, arch, method)
        assert bool([x for x in w if issubclass(x.category, nb.PerformanceWarning)]) == expect


# ---------------------------------------------------------------- input validation (reference fields.py:485-565)
def test_validate_inputs():
    pos = np.zeros((5, 3))
    v = fields._validate_inputs(pos, 2.0, 0.1, "float32_kahan", "Spline")
    assert v["N"] == 5 and v["mass_scalar"] == 2.0 and v["h_scalar"] == 0.1 and v["kernel_id"] == 4
    assert v["dtype_np"] == np.float32 and v["pos_cpu"].dtype == np.float32
    v = fields._validate_inputs(pos, np.ones(5), np.ones(5), "float64", "plummer")
    assert v["mass_cpu"].dtype == np.float64 and v["kernel_id"] == 1
    for bad in [dict(pos=np.zeros((5, 2))), dict(mass=np.ones(4)), dict(softening=np.ones(6)),
                dict(precision="f16"), dict(kernel="cubic")]:
        kw = dict(pos=pos, mass=1.0, softening=0.0, precision="float32", kernel="spline")
        kw.update(bad)
        with pytest.raises(ValueError):
            fields._validate_inputs(**kw)
    assert fields.KERNEL_MAP == {"newtonian": 0, "plummer": 1, "dehnen_k1": 2, "dehnen_k2": 3, "spline": 4}
    assert fields.G_DEFAULT == 4.300917270069976e-06


def test_gpu_calls_fail_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(ImportError):
        nb.compute_nbody_forces_gpu(np.zeros((4, 3)), 1.0)
    with pytest.raises(ImportError):
        nb.run_nbody_gpu(np.zeros((4, 6)), np.ones(4), 0, 1, 0.1, 0.1, save_snapshots=False)
    assert nb.get_gpu_info() == {"available": False}


# ---------------------------------------------------------------- make_plummer_sphere (reference run.py:1225-1325)
def _sha16(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_plummer_generator_bit_identical_to_reference(golden):
    g = golden("plummer_ic.npz")
    xv, m = nb.make_plummer_sphere(1000)
    assert np.array_equal(xv, g["xv_n1000_default"]) and m[0] == float(g["m0_n1000_default"])
    xv, m = nb.make_plummer_sphere(2000, M_total=1e7, a=0.3, seed=42070)
    assert np.array_equal(xv, g["xv_n2000_c2star"])
    xv, m = nb.make_plummer_sphere(10_000, M_total=1e9, a=1.0, seed=42069)
    assert _sha16(xv) == str(g["sha_n10000_c1"]) == "ea0c6ac8c7ccc56e"
    assert np.array_equal(xv[:64], g["xv_head_n10000_c1"]) and np.array_equal(xv[-64:], g["xv_tail_n10000_c1"])
    assert m.shape == (10_000,) and m.dtype == np.float64 and np.all(m == 1e5)


def test_plummer_generator_physics():
    xv, m = nb.make_plummer_sphere(50_000, M_total=1e9, a=1.0, seed=1)
    assert np.abs(xv.mean(axis=0)).max() < 1e-10                    # COM-centred
    r = np.linalg.norm(xv[:, :3], axis=1)
    assert abs(np.median(r) - 1.3048) < 0.03                          # half-mass radius 1.3048 a
    ke = 0.5 * np.sum(m * np.sum(xv[:, 3:] ** 2, axis=1))
    pe = -(3 * np.pi / 32) * nb.G_DEFAULT * 1e9 ** 2 / 1.0
    assert abs(ke / abs(pe) - 0.5) < 0.02                             # virial equilibrium


# ---------------------------------------------------------------- restart / snapshot.times (reference nbody_io.py)
def test_restart_roundtrip(tmp_path):
    xv = np.random.default_rng(0).standard_normal((7, 6))
    nbody_io._save_restart(xv, 1.5, 42, tmp_path, 3, mass_arr=np.ones(7), softening_arr=np.full(7, 0.1),
                           species_names=["dark", "bh"], species_N=[6, 1])
    ps, t, step, cnt, m, h, names, ns = nbody_io._load_restart(tmp_path)
    assert np.array_equal(ps, xv) and (t, step, cnt) == (1.5, 42, 3)
    assert names == ["dark", "bh"] and ns == [6, 1] and m.dtype == np.float64 and np.all(h == 0.1)
    raw = np.load(tmp_path / "restart.npz")
    assert set(raw.files) == {"phase_space", "time", "step", "snapshot_counter", "mass_arr", "softening_arr",
                              "species_names", "species_N"}
    nbody_io._save_restart(xv, 0.0, 0, tmp_path / "legacy", 0)
    assert nbody_io._load_restart(tmp_path / "legacy")[4:] == (None, None, None, None)
    assert nbody_io._load_restart(tmp_path / "missing") is None


def test_snapshot_times_file(tmp_path):
    for idx, t in [(2, 0.2), (0, 0.0), (1, 0.1), (1, 0.15)]:
        nbody_io._update_snapshot_times(tmp_path, idx, t)
    arr = np.loadtxt(tmp_path / "snapshot.times", comments="#")
    assert np.array_equal(arr[:, 0], [0, 1, 2]) and np.allclose(arr[:, 1], [0.0, 0.15, 0.2])
    assert (tmp_path / "snapshot.times").read_text().startswith("# snap_index time")


def test_snapshot_filename_distribution(tmp_path):
    f = nbody_io._snapshot_filename
    assert f(tmp_path, 5, None, None, None).name == "snapshot.h5"
    assert f(tmp_path, 5, None, 1, 10).name == "snapshot.h5"
    assert f(tmp_path, 0, None, 4, 10).name == "snapshot.000.h5"
    assert f(tmp_path, 9, None, 4, 10).name == "snapshot.003.h5"
    assert f(tmp_path, 5, None, 4, None).name == "snapshot.001.h5"


# ---------------------------------------------------------------- shard layout
@pytest.mark.parametrize("N,world", [(1, 1), (1000, 8), (1_000_000, 8), (262_144, 8), (2_000_001, 8), (257, 2),
                                     (120_000, 4), (255, 3)])
def test_shard_layout(N, world):
    L = ShardLayout(N, world)
    assert L.shard % L.tile == 0 and L.n_slots == L.shard * world and L.n_slots >= N
    counts = [L.count(r) for r in range(world)]
    assert sum(counts) == N and all(0 <= c <= L.shard for c in counts)
    covered = np.zeros(N, dtype=int)
    for r in range(world):
        covered[L.rows(r)] += 1
        t0, t1 = L.tile_range(r)
        assert t0 * L.tile == L.first(r) and (t1 - t0) == L.tiles_per_rank
        assert L.first(r) % L.tile == 0
    assert np.all(covered == 1)
    # only the tail of the global range is padding: slot index == particle index
    full = [r for r in range(world) if counts[r] == L.shard]
    assert full == list(range(len(full)))


# ---------------------------------------------------------------- C ABI surface (no compute calls)
def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "nbody_sm100.h").read_text()
    declared = set(re.findall(r"\b(nb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/nbody_sm100.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) <= declared


def test_integration_guide_covers_every_entry_point():
    """INTEGRATION.md shows the reference-side binding; every symbol of the header must be accounted for there."""
    header = (ROOT / "include" / "nbody_sm100.h").read_text()
    guide = (ROOT / "INTEGRATION.md").read_text()
    for name in sorted(set(re.findall(r"\b(nb_[a-z0-9_]+)\s*\(", header))):
        assert name in guide, f"{name} is declared in include/nbody_sm100.h but not mentioned in INTEGRATION.md"


def test_library_host_side_queries():
    lib = _lib.load()
    assert lib.nb_version() >= 100
    assert lib.nb_tile_j() == 256
    assert lib.nb_padded_count(1) == 256 and lib.nb_padded_count(256) == 256 and lib.nb_padded_count(257) == 512
    assert lib.nb_packed_bytes(_lib.PREC_F32, 1000) == 4 * (256 * 20 + 16)
    assert lib.nb_packed_bytes(_lib.PREC_F64, 1000) == 4 * (256 * 40 + 16)
    assert lib.nb_block_i(_lib.PREC_F32) % 256 == 0 and lib.nb_block_i(_lib.PREC_F64) % 128 == 0
    assert lib.nb_workspace_bytes(_lib.PREC_F32, 1000, 592, 1) > 0
    assert lib.nb_workspace_bytes(_lib.PREC_F32_COMP, 1000, 592, 2) == 2 * lib.nb_workspace_bytes(_lib.PREC_F32_COMP, 1000, 592, 1)
    assert lib.nb_error_string(0) == b"ok" and b"invalid" in lib.nb_error_string(-1)
    assert lib.nb_set_variant(0, 99) == -1 and lib.nb_set_variant(0, 0) == 0 and lib.nb_get_variant(0) == 0
    assert lib.nb_set_variant(0, -2) == -1 and lib.nb_set_variant(0, -1) == 0 and lib.nb_get_variant(0) == -1   # automatic
    assert lib.nb_energy_workspace_bytes(10_000) >= 8 * 8
    # argument validation happens before any CUDA call
    assert lib.nb_pack(7, 0, None, 0, None, 0, None, 0, 0.0, None, 0, 0.0, 0, 0, 0.0, None) == -1
    assert lib.nb_sweep(0, 9, 1, None, 0, 0, 0, 0, 0, 0, 0, 0.0, None, 0, 0, 1, None) == -1


def test_product_code_never_touches_the_oracle():
    for f in (ROOT / "nbody_streams_b200").rglob("*.py"):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, f
    assert "oracle" not in (ROOT / "include" / "nbody_sm100.h").read_text()
