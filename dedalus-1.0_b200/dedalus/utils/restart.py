"""Restart a run from a snapshot directory (reference: dedalus/utils/restart.py:33-99).

A snapshot (TimeStepBase.snapshot, time_step.py:112-151) holds, per rank, the pickled physics / state-layout /
integrator objects, the source of the forcing functions and the field arrays -- HDF5 with the reference's layout when
h5py is importable, otherwise .npy files with a JSON index.  `restart` rebuilds the three objects, re-attaches the
forcing functions by name and copies the saved arrays into the state's device buffers."""
import importlib.util
import inspect
import json
import os
import pickle
import time

import numpy as np
import torch

from .parallelism import com_sys

OBJECT_FILENAME = "dedalus_obj_%04i.cpkl"


def restart(snap_dir):
    """(RHS, data, ti) as they were when `snap_dir` was written."""
    rank = com_sys.myproc
    with open(os.path.join(snap_dir, OBJECT_FILENAME % rank), "rb") as f:
        RHS = pickle.load(f)
        data = pickle.load(f)
        ti = pickle.load(f)
    fn = os.path.join(snap_dir, "forcing_functions.py")
    forcing = None
    if os.path.exists(fn) and os.path.getsize(fn) > 0:
        spec = importlib.util.spec_from_file_location("forcing_functions", fn)
        forcing = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(forcing)
    for k, func_name in RHS._forcing_function_names.items():
        for name, func in (inspect.getmembers(forcing) if forcing is not None else ()):
            if name == func_name:
                RHS.forcing_functions[k] = func
        if k not in RHS.forcing_functions:
            raise ValueError("Failed to load driver function for %s. Snapshot may be corrupted." % k)
    load_all_data(data, snap_dir)
    ti.RHS = RHS
    ti._start_time = time.time()
    return RHS, data, ti


def load_all_data(data, location):
    """Fill every component of `data` from a snapshot: `location` is the snapshot directory, or the HDF5 file of this
    rank as in the reference (restart.py:72-99)."""
    rank = com_sys.myproc
    if os.path.isdir(location):
        h5 = os.path.join(location, "data.cpu%04i" % rank)
        index = os.path.join(location, "fields.cpu%04i.json" % rank)
    else:
        h5, index = location, None
    if index is not None and os.path.exists(index) and not os.path.exists(h5):
        with open(index) as f:
            meta = json.load(f)
        for name, i, c in data.components():
            try:
                entry = meta["fields"]["%s/%i" % (name, i)]
                arr = np.load(os.path.join(location, entry["file"]))
            except (KeyError, OSError):
                raise KeyError("Data File field %s component %i is missing or corrupt" % (name, i))
            c[entry["space"]] = torch.from_numpy(arr)
        return                      # data.time travels in the pickle (the file's /time is the integrator's clock)
    import h5py
    with h5py.File(h5, mode="r") as f:
        for name, i, c in data.components():
            key = "/fields/%s/%i" % (name, i)
            try:
                dset = f[key]
                space = dset.attrs["space"]
            except KeyError:
                raise KeyError("Data File field %s component %i is missing or corrupt" % (name, i))
            if isinstance(space, bytes):
                space = space.decode()
            c[space] = torch.from_numpy(np.asarray(dset[...]))


def identify_version(snap_dir, cpu_file="data.cpu0000"):
    """The version string of the code that wrote a snapshot (restart.py:101-110): attribute `hg_version` of the HDF5 field
    file; for the .npy + JSON fallback TimeStepBase.snapshot writes without h5py, the same value from the JSON index."""
    filename = os.path.join(snap_dir, cpu_file)
    if os.path.exists(filename):
        import h5py
        with h5py.File(filename, mode="r") as fi:
            version = fi.attrs["hg_version"]
        return version.decode() if isinstance(version, bytes) else version
    index = os.path.join(snap_dir, cpu_file.replace("data.", "fields.") + ".json")
    with open(index) as f:
        return json.load(f).get("hg_version")
