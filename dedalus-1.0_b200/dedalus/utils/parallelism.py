"""Communication singleton and reductions (reference: dedalus/utils/parallelism.py:29-173).

The reference wraps mpi4py's COMM_WORLD; here one process drives one GPU and the process group
is torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests).  With no process
group initialised the system degrades to a single rank exactly as the reference does without
mpi4py (parallelism.py:33-50).
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except ImportError:  # pragma: no cover
    torch = None
    dist = None


class CommunicationSystem(object):
    """`com_sys`: rank / size bookkeeping.  `comm` is the torch.distributed module when a
    process group exists, else None (the reference's `comm is None` tests keep working)."""

    @property
    def comm(self):
        return dist if (dist is not None and dist.is_available() and dist.is_initialized()) else None

    @property
    def myproc(self):
        return dist.get_rank() if self.comm else 0

    @property
    def nproc(self):
        return dist.get_world_size() if self.comm else 1

    MPI = None
    _host_group = None

    def host_or(self, bits):
        """Bitwise OR of a small non-negative integer over all ranks, on the HOST (no device work, no stream
        synchronisation): a gloo group beside the NCCL one, created at the first call -- which is collective, like
        every later one.  Used where ranks must agree on what they know about caller-written buffers before any of
        them enters a device collective (Physics.sync_knowledge)."""
        if self.comm is None or self.nproc == 1:
            return int(bits)
        if self._host_group is None:
            self._host_group = dist.group.WORLD if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")
        t = torch.tensor([int(bits)], dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.BOR, group=self._host_group)
        return int(t.item())


com_sys = CommunicationSystem()


def _as_tensor(x):
    if torch.is_tensor(x):
        return x
    return torch.as_tensor(np.asarray(x))


def _all_reduce(value, op):
    """value: 0-d tensor on the device the process group expects."""
    if com_sys.comm is None:
        return value
    t = value.clone()
    if dist.get_backend() == "nccl" and not t.is_cuda:
        t = t.cuda()
    dist.all_reduce(t, op=op)
    return t


def _finish(t, reduce_all):
    """The reference returns the value on every rank for reduce_all, else only on rank 0."""
    out = t.item()
    if com_sys.comm is None or reduce_all or com_sys.myproc == 0:
        return out
    return None


def reduce_mean(data):
    data = _as_tensor(data)
    if com_sys.comm is None:
        return data.mean().item()
    total = _all_reduce(data.sum(), dist.ReduceOp.SUM)
    count = _all_reduce(torch.tensor(float(data.numel()), dtype=total.dtype, device=total.device), dist.ReduceOp.SUM)
    return _finish(total / count, False)


def reduce_sum(data, reduce_all=False):
    data = _as_tensor(data)
    return _finish(_all_reduce(data.sum(), dist.ReduceOp.SUM if dist else None), reduce_all)


def reduce_min(data, reduce_all=False):
    data = _as_tensor(data)
    return _finish(_all_reduce(data.min(), dist.ReduceOp.MIN if dist else None), reduce_all)


def reduce_max(data, reduce_all=False):
    data = _as_tensor(data)
    return _finish(_all_reduce(data.max(), dist.ReduceOp.MAX if dist else None), reduce_all)


def swap_indices(arr):
    """Exchange entries [0] and [1] (x-space order <-> FFTW-transposed k-space order)."""
    if isinstance(arr, np.ndarray):
        out = arr.copy()
    elif isinstance(arr, list):
        out = list(arr)
    else:
        raise NotImplementedError("swap_indices only implemented for numpy arrays and lists.")
    out[0], out[1] = out[1], out[0]
    return out


def load_all(field, snap_dir):
    """Concatenate every rank's part of one field component of a snapshot into a single array for analysis (reference:
    parallelism.py:54-100): `field` like 'u/0' or '/fields/u/0'; returns (data, space) with the reference's conventions --
    parts joined along axis 0, then axes 0 and 1 swapped (3-D).  Reads the HDF5 files when h5py is importable and they
    exist, else the .npy + JSON fallback TimeStepBase.snapshot writes (block ky ownership)."""
    import glob
    import json
    import os
    name = field[len("/fields/"):] if field.startswith("/fields/") else field
    h5 = sorted(glob.glob(os.path.join(snap_dir, "data.cpu*")))
    parts, space = [], None
    if h5:
        import h5py
        for fn in h5:
            with h5py.File(fn, "r") as fi:
                dset = fi["/fields/" + name]
                space = dset.attrs["space"]
                parts.append(np.asarray(dset[...]))
        if isinstance(space, bytes):
            space = space.decode()
    else:
        for fn in sorted(glob.glob(os.path.join(snap_dir, "fields.cpu*.json"))):
            with open(fn) as f:
                meta = json.load(f)
            entry = meta["fields"][name]
            space = entry["space"]
            parts.append(np.load(os.path.join(snap_dir, entry["file"])))
    if not parts:
        raise IOError("no snapshot data in %s" % snap_dir)
    data = np.concatenate(parts, axis=0)
    if data.ndim == 3:
        data = np.transpose(data, axes=[1, 0, 2])
    return data, space
