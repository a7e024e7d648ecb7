"""Host-side plan cache: wavenumbers, dealias masks and the ddl_plan handle shared by every
component of the same (shape, length, dealiasing).  Replaces fftw.create_data / fftw.rPlan
(dedalus/utils/fftw/_fftw.pyx:81-185,246-309) and the per-representation k / mask setup
(representations.py:204-233,359-417): one plan, no per-component scratch."""
import ctypes as C

import numpy as np
import numpy.fft as npfft
import torch

from .._lib import lib, check
from ..utils.parallelism import swap_indices, com_sys

_PLANS = {}


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("dedalus (B200) needs a CUDA device: there is no CPU path.")
    return torch.device("cuda", torch.cuda.current_device())


def current_stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def wavenumbers(shape, length):
    """k arrays exactly as representations.py:204-229 builds them (same float expressions):
    returns (global kspace shape, ktrans, dk, kny, {name: 1-D float64 array})."""
    ndim = len(shape)
    xs = np.array(shape)
    length = np.asarray(length, dtype=float)
    ks = xs.copy()
    ks[-1] = ks[-1] // 2 + 1
    kshape = np.array(swap_indices(ks))
    if ndim == 2:
        ktrans = {"x": 0, 0: "x", "y": 1, 1: "y"}
    else:
        ktrans = {"x": 2, 2: "x", "y": 0, 0: "y", "z": 1, 1: "z"}
    dk = swap_indices(2 * np.pi / length)
    kny = swap_indices(np.pi * xs / length)
    xs_sw = swap_indices(xs)
    k = {}
    for i, ksize in enumerate(kshape):
        xsize = xs_sw[i]
        if i == ktrans["x"]:
            ki = npfft.fftfreq(xsize)[:ksize] * 2.0 * kny[i]
            if xsize % 2 == 0:
                ki[-1] *= -1.0
        else:
            ki = npfft.fftfreq(ksize) * 2.0 * kny[i]
            if xsize % 2 == 0:
                ki[ksize // 2] *= -1.0
        k[ktrans[i]] = np.ascontiguousarray(ki)
    return kshape, ktrans, dk, kny, k


def keep_masks(dealiasing, ktrans, kny, k):
    """Per-axis 'mode survives' masks with the reference's own comparisons
    (dealias_cy_3d.pyx:34-36: zero if |k| >= 2/3 k_nyquist; representations.py:442-455)."""
    keep = {}
    for name, kv in k.items():
        kn = kny[ktrans[name]]
        if dealiasing in ("2/3", "2/3 cython"):
            keep[name] = ~((kv >= 2.0 / 3.0 * kn) | (kv <= -2.0 / 3.0 * kn))
        elif dealiasing in ("None", None, 0):
            keep[name] = ~(np.abs(kv) == kn)
        elif dealiasing == "2/3 spherical":
            # representations.py:410-417: zero where |k| >= 2/3 min(kny).  The plan prunes to the cube that bounds this
            # sphere; the representation applies the sphere itself (FourierRepresentation._sphere)
            keep[name] = ~(np.abs(kv) >= 2.0 / 3.0 * np.min(kny))
        else:
            raise NotImplementedError("Specified dealiasing method not implemented.")
    return keep


class Plan(object):
    """nranks > 1: this rank's part of a slab decomposition (3-D only).  x-space is split along
    z, k-space along ky, as FFTW-MPI does for the reference (_fftw.pyx:114-148,
    representations.py:180-186,231-233); `k['y']`, `kshape_local`, `xshape_local` and the
    offsets describe the local slab, the `*_np` arrays stay global."""

    def __init__(self, shape, length, dealiasing, nranks=1, rank=0, ky_layout="block", full_ky=False):
        self.shape = tuple(int(s) for s in shape)
        self.ndim = len(self.shape)
        self.length = tuple(float(x) for x in length)
        self.dealiasing = dealiasing
        self.nranks, self.rank = int(nranks), int(rank)
        if ky_layout not in ("block", "cyclic"):
            raise ValueError("parallel.ky_layout must be 'block' or 'cyclic'")
        self.ky_layout = ky_layout if self.nranks > 1 else "block"
        self.kshape, self.ktrans, self.dk, self.kny, self.k_np = wavenumbers(self.shape, self.length)
        self.keep_np = keep_masks(dealiasing, self.ktrans, self.kny, self.k_np)
        # shearing box: the ky mask depends on kx and time and is applied by the representation
        # (representations.py:627-642); the plan keeps every ky row, the passes prune nothing along y
        self.full_ky = bool(full_ky)
        if self.full_ky:
            self.keep_np["y"] = np.ones_like(self.keep_np["y"])
        self.device = device()
        shp = np.array(self.shape, dtype=np.int64)
        keep8 = {n: np.ascontiguousarray(v.astype(np.uint8)) for n, v in self.keep_np.items()}
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self.handle = C.c_void_p()
        if self.nranks > 1 and self.ndim != 3:
            raise NotImplementedError("Slab decomposition is 3-D only: 2-D grids fit one GPU and run as replicas.")
        check(lib.ddl_plan_create_slab(C.byref(self.handle), self.ndim, vp(shp), vp(self.k_np["x"]), vp(self.k_np["y"]),
                                       vp(self.k_np["z"]) if self.ndim == 3 else None, vp(keep8["x"]), vp(keep8["y"]),
                                       vp(keep8["z"]) if self.ndim == 3 else None, self.nranks, self.rank,
                                       1 if self.ky_layout == "cyclic" else 0))
        self.kshape_local = self.kshape.copy()
        self.xshape_local = np.array(self.shape)
        self.koffset = self.xoffset = 0
        self.krows = np.arange(int(self.kshape[0]))          # global ky index of every local k-space row
        if self.nranks > 1:
            self.kshape_local[0] = self.kshape[0] // self.nranks
            self.xshape_local[0] = self.shape[0] // self.nranks
            self.xoffset = self.rank * int(self.xshape_local[0])
            if self.ky_layout == "cyclic":
                self.koffset = self.rank
                self.krows = np.arange(self.rank, int(self.kshape[0]), self.nranks)
            else:
                self.koffset = self.rank * int(self.kshape_local[0])
                self.krows = np.arange(self.koffset, self.koffset + int(self.kshape_local[0]))
        # broadcast-shaped device copies for the Python-level API (comp.k['x'] etc.); axis 0 of
        # k-space is restricted to the local slab (representations.py:232-233)
        self.k = {}
        for name, kv in self.k_np.items():
            i = self.ktrans[name]
            if i == 0 and self.nranks > 1:
                kv = kv[self.krows]
            shp_b = [1] * self.ndim
            shp_b[i] = len(kv)
            self.k[name] = torch.from_numpy(np.ascontiguousarray(kv)).to(self.device).reshape(shp_b)
        self._work = None
        self._pipe = None
        self._boxes = None
        self.nmodes = int(np.prod(self.kshape))

    @property
    def pipeline(self):
        """The slab pipeline of this rank (dedalus/data_objects/slab.py); 3-D plans only."""
        if self._pipe is None:
            from .slab import SlabPipeline
            import os
            from ..config import decfg
            kind = os.environ.get("DEDALUS_SLAB_EXCHANGE", decfg.get("parallel", "exchange"))
            if self.full_ky or (kind == "peer" and any(n < 16 or (n & (n - 1)) for n in self.shape[:2])):
                # the peer-store passes exist in the specialised strided kernels only (powers of two >= 16);
                # any other nz / ny exchanges through the collective
                kind = "collective"
            self._pipe = SlabPipeline(lib, self.handle, self.device, stream=current_stream, exchange=kind)
        return self._pipe

    def workspace(self, nbytes):
        if self._work is None or self._work.numel() < nbytes:
            self._work = None
            self._work = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._work

    def retained_boxes(self):
        """(shape3, boxes) of the local k-space array for ddl_copy_boxes: the index boxes the plan's mask retains (runs of kept
        indices along every axis; two per full axis, one along the half axis), int64 arrays.  2-D arrays get a leading axis of 1."""
        if self._boxes is None:
            runs = []
            for axis in range(self.ndim):
                name = [n for n, i in self.ktrans.items() if isinstance(n, str) and i == axis][0]
                keep = np.asarray(self.keep_np[name]).astype(bool).ravel()
                if axis == 0 and self.nranks > 1:
                    keep = keep[self.krows]
                edges = np.flatnonzero(np.diff(np.concatenate(([0], keep.astype(np.int8), [0]))))
                runs.append(list(zip(edges[::2].tolist(), edges[1::2].tolist())))
            runs = [[(0, 1)]] * (3 - self.ndim) + runs
            boxes = [(a0, b0, a1, b1, a2, b2) for a0, b0 in runs[0] for a1, b1 in runs[1] for a2, b2 in runs[2]]
            shape3 = np.array([1] * (3 - self.ndim) + [int(n) for n in self.kshape_local], dtype=np.int64)
            self._boxes = (shape3, np.ascontiguousarray(np.array(boxes, dtype=np.int64).reshape(-1, 6)))
        return self._boxes

    def transform_workspace(self):
        return self.workspace(lib.ddl_workspace_bytes(self.handle, 1, 1))

    def rhs_workspace(self, physics_id):
        return self.workspace(lib.ddl_rhs_workspace_bytes(self.handle, physics_id))

    def __del__(self):
        try:
            if self.handle:
                lib.ddl_plan_destroy(self.handle)
        except Exception:
            pass


def get_plan(shape, length, dealiasing, full_ky=False):
    """One plan per (grid, dealiasing, device, process-group size): with a torch.distributed
    process group of P > 1 ranks a 3-D grid is slab-decomposed over it."""
    import os
    from ..config import decfg
    nranks, rank = com_sys.nproc, com_sys.myproc
    layout = os.environ.get("DEDALUS_KY_LAYOUT", decfg.get("parallel", "ky_layout")) if nranks > 1 else "block"
    key = (tuple(int(s) for s in shape), tuple(float(x) for x in length), str(dealiasing),
           torch.cuda.current_device() if torch.cuda.is_available() else -1, nranks, rank, layout, bool(full_ky))
    pl = _PLANS.get(key)
    if pl is None:
        pl = _PLANS[key] = Plan(shape, length, dealiasing, nranks, rank, layout, full_ky=full_ky)
    return pl
