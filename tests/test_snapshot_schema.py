"""HDF5 snapshot schema without h5py: this package's writer and the REFERENCE's writer both ran
through tests/fake_h5py.py; the reference's tree is committed as tests/golden/snapshots_schema.npz
(tests/golden/make_golden_snapshots.py, which also checked that the reference's ParticleReader reads this package's
files).  Here: our tree == the reference's tree, entry by entry, dtype and values."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import fake_h5py  # noqa: E402

GOLD = np.load(HERE / "golden" / "snapshots_schema.npz")


def _golden_script():
    spec = importlib.util.spec_from_file_location("make_golden_snapshots", HERE / "golden" / "make_golden_snapshots.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture
def h5(monkeypatch):
    monkeypatch.setitem(sys.modules, "h5py", fake_h5py)
    return fake_h5py


def test_writer_tree_equals_the_references(h5, tmp_path):
    from nbody_streams_b200 import nbody_io
    from nbody_streams_b200.species import Species

    g = _golden_script()
    g.write_all(nbody_io, Species, tmp_path)
    ours = g.collect(tmp_path)
    assert set(ours) == set(GOLD.files)
    for k in GOLD.files:
        assert ours[k].dtype == GOLD[k].dtype, k
        assert np.array_equal(ours[k], GOLD[k]), k
    # a few entries spelled out, so the schema is readable here
    assert "multi/snapshot.h5|/snapshots/snap.002" in ours and "multi/snapshot.h5|/snapshots@snap_time.002" in ours
    assert int(ours["multi/snapshot.h5|/properties@n_species"]) == 3
    assert [n.decode() for n in ours["multi/snapshot.h5|/properties@species_names"]] == ["dark", "star", "bh"]
    assert float(ours["multi/snapshot.h5|/properties/dark/m"]) == 1e4                 # uniform -> scalar dataset
    assert ours["multi/snapshot.h5|/properties/star/m_array"].shape == (14,)          # varying -> compressed array
    assert str(ours["multi/snapshot.h5|/snapshots/snap.000#compression"]) == "gzip"
    assert {k.split("|")[0] for k in ours if k.startswith("split/") and k.endswith(".h5|/properties/time_step")} == \
        {"split/snapshot.000.h5", "split/snapshot.001.h5", "split/snapshot.002.h5"}
    assert "snap_index time" in str(ours["multi/snapshot.times|text"])


def test_existing_snapshot_is_never_overwritten(h5, tmp_path):
    from nbody_streams_b200 import nbody_io

    a, b = np.zeros((3, 6)), np.ones((3, 6))
    nbody_io._save_snapshot(a, 0, 0.0, tmp_path, mass_dark=1.0)
    nbody_io._save_snapshot(b, 0, 9.0, tmp_path, mass_dark=1.0)
    flat = h5.flatten(tmp_path / "snapshot.h5")
    assert np.array_equal(flat["/snapshots/snap.000"], a) and float(flat["/snapshots@snap_time.000"]) == 0.0


def test_nearly_uniform_masses_follow_the_references_tolerance(h5, tmp_path):
    from nbody_streams_b200 import nbody_io
    from nbody_streams_b200.species import Species

    m = np.full(5, 2.0)
    m[3] *= 1 + 1e-12                       # inside the reference's rtol 1e-10 -> stored as a scalar
    nbody_io._save_snapshot(np.zeros((5, 6)), 0, 0.0, tmp_path / "u", species=[Species("a", 5, m, softening=0.1)])
    m2 = m.copy()
    m2[3] *= 1 + 1e-8                       # outside -> stored as an array
    nbody_io._save_snapshot(np.zeros((5, 6)), 0, 0.0, tmp_path / "v", species=[Species("a", 5, m2, softening=0.1)])
    assert "/properties/a/m" in h5.flatten(tmp_path / "u" / "snapshot.h5")
    assert "/properties/a/m_array" in h5.flatten(tmp_path / "v" / "snapshot.h5")
