"""A stand-in for the part of the h5py API the snapshot code uses (TEST INFRASTRUCTURE; this image has no h5py and no HDF5
library): File(path, mode) as a context manager, create_group / create_dataset, item access by path, `[...]` / `[:]` on
datasets, `.attrs` on files, groups and datasets.  Files are pickles, NOT HDF5: the stand-in lets the HDF5 BRANCH of
TimeStepBase.snapshot / restart / load_all / identify_version execute (argument order, attribute names, layout
/time, /fields/<name>/<comp>, attribute 'space' -- reference time_step.py:141-151, fields.py:118-125, restart.py:72-110);
it says nothing about the HDF5 byte format, which only a real h5py can check."""
import pickle

import numpy as np


class _Node(object):
    def __init__(self):
        self.attrs = {}


class Dataset(_Node):
    def __init__(self, shape=None, dtype=None, data=None):
        _Node.__init__(self)
        self.value = np.array(data) if data is not None else np.zeros(shape, dtype=dtype)
        self.shape, self.dtype = self.value.shape, self.value.dtype

    def __setitem__(self, key, v):
        self.value[key] = v

    def __getitem__(self, key):
        return self.value[key]

    def __array__(self, dtype=None, copy=None):
        return self.value if dtype is None else self.value.astype(dtype)


class Group(_Node):
    def __init__(self):
        _Node.__init__(self)
        self.items_ = {}

    def _walk(self, path, create=False):
        node = self
        for part in [p for p in path.split("/") if p]:
            if part not in node.items_:
                if not create:
                    raise KeyError(path)
                node.items_[part] = Group()
            node = node.items_[part]
        return node

    def create_group(self, name):
        return self._walk(name, create=True)

    def create_dataset(self, name, shape=None, dtype=None, data=None):
        parts = [p for p in name.split("/") if p]
        parent = self._walk("/".join(parts[:-1]), create=True)
        d = parent.items_[parts[-1]] = Dataset(shape, dtype, data)
        return d

    def __getitem__(self, path):
        return self._walk(path)

    def __contains__(self, path):
        try:
            self._walk(path)
            return True
        except KeyError:
            return False


class File(Group):
    def __init__(self, path, mode="r"):
        Group.__init__(self)
        self._path, self._mode = path, mode
        if mode in ("r", "r+", "a"):
            try:
                with open(path, "rb") as f:
                    root = pickle.load(f)
                self.items_, self.attrs = root.items_, root.attrs
            except FileNotFoundError:
                if mode != "a":
                    raise IOError("unable to open file " + path)

    def close(self):
        if self._mode != "r":
            root = Group()
            root.items_, root.attrs = self.items_, self.attrs
            with open(self._path, "wb") as f:
                pickle.dump(root, f)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False
