"""In-memory stand-in for the small part of h5py the snapshot writers / readers of this path use
(h5py is not installed in this image).  TEST INFRASTRUCTURE: lets tests/ exercise the HDF5 snapshot schema
— groups, datasets, attributes, append mode — and compare the tree written by this package with the tree
the REFERENCE's own writer produces through the same stand-in (tests/golden/make_golden_snapshots.py).

Supported: File(path, "a"|"r"|"r+"|"w") as context manager (the tree is pickled to `path` on close so that
append mode sees earlier writes), require_group, create_group, create_dataset(name, data=, compression=...),
`name in group`, group["a/b"], group.get, group.keys(), .attrs (dict), dataset[:] / [()] / [...] / .shape / .dtype.
"""
from __future__ import annotations

import pickle
from pathlib import Path

import numpy as np


class Dataset:
    def __init__(self, data, compression=None):
        self.data = np.array(data)
        self.compression = compression
        self.attrs = {}

    shape = property(lambda self: self.data.shape)
    dtype = property(lambda self: self.data.dtype)

    def __getitem__(self, key):
        return self.data[key]

    def __len__(self):
        return len(self.data)


class Group:
    def __init__(self):
        self.members = {}
        self.attrs = {}

    def _walk(self, path, create=False):
        node = self
        for part in [p for p in str(path).split("/") if p]:
            if part not in node.members:
                if not create:
                    raise KeyError(f"Unable to open object (object '{part}' doesn't exist)")
                node.members[part] = Group()
            node = node.members[part]
        return node

    def require_group(self, name):
        node = self._walk(name, create=True)
        if not isinstance(node, Group):
            raise TypeError(f"Incompatible object (Dataset) already exists at {name}")
        return node

    def create_group(self, name):
        if name in self:
            raise ValueError(f"Unable to create group (name already exists): {name}")
        return self._walk(name, create=True)

    def create_dataset(self, name, shape=None, dtype=None, data=None, compression=None, **_ignored):
        if name in self:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        parts = [p for p in str(name).split("/") if p]
        parent = self._walk("/".join(parts[:-1]), create=True)
        if data is None:
            data = np.zeros(shape, dtype=dtype)
        ds = Dataset(np.asarray(data, dtype=dtype) if dtype is not None else data, compression)
        parent.members[parts[-1]] = ds
        return ds

    def __contains__(self, name):
        try:
            self._walk(name)
            return True
        except KeyError:
            return False

    def __getitem__(self, name):
        return self._walk(name)

    def get(self, name, default=None):
        return self[name] if name in self else default

    def keys(self):
        return self.members.keys()

    def values(self):
        return self.members.values()

    def __iter__(self):
        return iter(self.members)

    def items(self):
        return self.members.items()


class File(Group):
    def __init__(self, name, mode="r", **_ignored):
        super().__init__()
        self.filename = str(name)
        self.mode = mode
        path = Path(self.filename)
        if mode in ("r", "r+") and not path.exists():
            raise FileNotFoundError(f"Unable to open file (unable to open file: name = '{name}')")
        if mode in ("r", "r+", "a") and path.exists():
            with open(path, "rb") as fh:
                self.members, self.attrs = pickle.load(fh)

    def close(self):
        if self.mode != "r":
            with open(self.filename, "wb") as fh:
                pickle.dump((self.members, self.attrs), fh)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def flatten(path) -> dict:
    """{'/group/dataset': array, '/group@attr': array, '/group/dataset#compression': str} of a file written above."""
    out = {}

    def rec(node, prefix):
        for k, v in node.attrs.items():
            out[f"{prefix or '/'}@{k}"] = np.asarray(v)
        if isinstance(node, Dataset):
            out[prefix] = node.data
            out[f"{prefix}#compression"] = np.asarray(str(node.compression))
            return
        for name, child in node.members.items():
            rec(child, f"{prefix}/{name}")

    with File(path, "r") as f:
        rec(f, "")
    return out
