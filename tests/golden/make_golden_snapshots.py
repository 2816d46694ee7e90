#!/usr/bin/env python3
"""Golden trees for the HDF5 snapshot schema: run the REFERENCE's own writer (nbody_io._save_snapshot,
_update_snapshot_times; reference nbody_io.py:770-929, :1036-1074) through tests/fake_h5py.py (h5py is not in
this image) and store the flattened trees in tests/golden/snapshots_schema.npz.  Also checks that the
reference's ParticleReader reads files written by THIS package's writer through the same stand-in.

Build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_snapshots.py
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types
import warnings
from pathlib import Path

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nb_numba_cache")
sys.dont_write_bytecode = True
import numpy as np  # noqa: E402

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))          # tests/ (fake_h5py)
sys.path.insert(0, str(HERE.parent.parent))   # repo root (nbody_streams_b200)
import fake_h5py  # noqa: E402

REF = Path("/root/reference")


def scenarios(Species):
    rng = np.random.default_rng(5)
    xv = rng.standard_normal((40, 6))
    sp = [Species.dark(25, mass=1e4, softening=0.05),
          Species.star(14, mass=np.linspace(400.0, 600.0, 14), softening=np.linspace(0.01, 0.03, 14)),
          Species("bh", 1, 1e8, softening=0.001)]
    return xv, sp


def write_all(mod, Species, root: Path):
    """The same calls through `mod` (the reference's nbody_io or this package's)."""
    xv, sp = scenarios(Species)
    a = root / "multi"
    for i, t in enumerate((0.0, 0.5, 1.0)):
        mod._save_snapshot(xv + i, i, t, a, species=sp, time_step=1e-3, num_files_to_write=1, total_expected_snapshots=3)
        mod._update_snapshot_times(a, i, t)
    mod._save_snapshot(xv + 99, 1, 0.5, a, species=sp, time_step=1e-3, num_files_to_write=1)      # never overwrites
    b = root / "legacy"
    for i, t in enumerate((0.0, 2.0)):
        mod._save_snapshot(xv * (i + 1), i, t, b, mass_dark=123.0, num_files_to_write=1, total_expected_snapshots=2)
        mod._update_snapshot_times(b, i, t)
    c = root / "split"
    for i in range(7):
        mod._save_snapshot(xv - i, i, 0.1 * i, c, num_dark=30, num_star=10, mass_dark=2.0, mass_star=1.0, eps_dark=0.05,
                           eps_star=0.01, time_step=1e-4, num_files_to_write=3, total_expected_snapshots=7)
        mod._update_snapshot_times(c, i, 0.1 * i)


def collect(root: Path) -> dict:
    out = {}
    for f in sorted(root.rglob("*.h5")):
        for k, v in fake_h5py.flatten(f).items():
            out[f"{f.relative_to(root)}|{k}"] = v
    for f in sorted(root.rglob("snapshot.times")):
        out[f"{f.relative_to(root)}|text"] = np.asarray(f.read_text())
    return out


def main():
    warnings.simplefilter("ignore")
    sys.modules["h5py"] = fake_h5py
    pkg = types.ModuleType("nbody_streams")
    pkg.__path__ = [str(REF / "nbody_streams")]
    sys.modules["nbody_streams"] = pkg
    ref_io = importlib.import_module("nbody_streams.nbody_io")
    ref_species = importlib.import_module("nbody_streams.species")
    with tempfile.TemporaryDirectory() as td:
        write_all(ref_io, ref_species.Species, Path(td))
        gold = collect(Path(td))
    np.savez_compressed(HERE / "snapshots_schema.npz", **gold)
    print(f"snapshots_schema.npz: {len(gold)} entries, {(HERE / 'snapshots_schema.npz').stat().st_size / 1024:.1f} KiB")

    # the reference's reader on files written by THIS package's writer
    from nbody_streams_b200 import nbody_io as our_io
    from nbody_streams_b200.species import Species as OurSpecies
    with tempfile.TemporaryDirectory() as td:
        write_all(our_io, OurSpecies, Path(td))
        ours = collect(Path(td))
        assert set(ours) == set(gold), set(ours) ^ set(gold)
        for k in gold:
            assert ours[k].dtype == gold[k].dtype and np.array_equal(ours[k], gold[k]), k
        xv, sp = scenarios(OurSpecies)
        reader = ref_io.ParticleReader(str(Path(td) / "multi" / "snapshot.h5"))
        part = reader.read_snapshot(2)
        assert np.array_equal(part.dark["posvel"], (xv + 2)[:25]) and np.array_equal(part.star["posvel"], (xv + 2)[25:39])
        assert np.array_equal(part.species["bh"]["posvel"], (xv + 2)[39:]) and part.time == 1.0
        assert np.allclose(part.star["mass"], np.linspace(400.0, 600.0, 14)) and float(part.dark["mass"][0]) == 1e4
        split = ref_io.ParticleReader(str(Path(td) / "split" / "snapshot.*.h5"))
        p5 = split.read_snapshot(5)
        assert np.array_equal(p5.dark["posvel"], (xv - 5)[:30]) and np.array_equal(p5.star["posvel"], (xv - 5)[30:])
        print("reference ParticleReader reads this package's files (multi-species single file, legacy split files)")
    print("this package's writer produces the reference's tree exactly")


if __name__ == "__main__":
    main()
