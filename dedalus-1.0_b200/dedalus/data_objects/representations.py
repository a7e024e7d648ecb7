"""Field-component representations (reference: dedalus/data_objects/representations.py).

`FourierRepresentation` keeps the reference's interface -- ``comp['kspace']`` / ``comp['xspace']``
lazy state machine, ``forward`` / ``backward``, ``deriv``, ``k2``, ``dealias`` and friends -- with
torch CUDA tensors as storage and the transforms done by the hand-written CUDA kernels behind
include/ddl.h.  k-space is transposed exactly like FFTW-MPI's output (representations.py:84-93):
3-D x-space (z,y,x) <-> k-space (ky,kz,kx); 2-D (y,x) <-> (kx,ky).
"""
import ctypes as C

import numpy as np
import torch

from .._lib import lib, check
from ..config import decfg
from ..utils.timer import timer
from . import plan as _plan


class Representation(object):
    """Base class: stores data and provides spatial derivatives (representations.py:47-55)."""

    def __init__(self, sd, shape, length):
        pass


class FourierRepresentation(Representation):
    """Component that is periodic (Fourier) in every direction."""

    timer = timer
    _static_k = True

    def __init__(self, sd, shape, length):
        self.sd = sd
        self.ndim = len(shape)
        if self.ndim not in (2, 3):
            raise ValueError("Must use either 2 or 3 dimensions.")
        if len(shape) != len(length):
            raise ValueError("Shape and Length must have same dimensions.")
        self.global_shape = {"xspace": np.array(shape)}
        self.length = np.asarray(length, dtype=float)
        self.dtype = {"kspace": "complex128", "xspace": "float64"}
        self._eps = {"kspace": np.finfo(np.complex128).eps, "xspace": np.finfo(np.float64).eps}

        method = decfg.get("FFT", "method")
        dealiasing = decfg.get("FFT", "dealiasing")
        if method not in ("cuda", "fftw", "numpy"):
            raise NotImplementedError("Specified FFT method not implemented.")

        if self.ndim == 2:
            self.xtrans = {"x": 1, 1: "x", "y": 0, 0: "y"}
        else:
            self.xtrans = {"x": 2, 2: "x", "y": 1, 1: "y", "z": 0, 0: "z"}

        self._plan = _plan.get_plan(shape, self.length, dealiasing, full_ky=not self._static_k)
        pl = self._plan
        self.ktrans = pl.ktrans
        self.global_shape["kspace"] = pl.kshape.copy()
        self.local_shape = {"kspace": pl.kshape_local.copy(), "xspace": pl.xshape_local.copy()}
        self.offset = {"xspace": pl.xoffset, "kspace": pl.koffset}
        # global ky index of every local k-space row (block slabs: offset + arange; cyclic: rank::P)
        self.local_rows = {"kspace": pl.krows}
        self.dk, self.kny, self.k = pl.dk, pl.kny, pl.k
        self.set_dealiasing(dealiasing)

        self._k = torch.zeros(tuple(int(n) for n in pl.kshape_local), dtype=torch.complex128, device=pl.device)
        self._xdata = None          # allocated on first use: most components never leave k-space
        # True while the spectrum is KNOWN to vanish outside the dealias mask (set by our own
        # kernels, cleared whenever the buffer is handed to the caller, who may write to it);
        # lets the RHS skip mask passes and the RK sweep visit the retained modes only
        self._clean = True
        self._checked = False       # _clean is False and a device check has confirmed modes outside the mask
        # True while the kx = 0 plane is KNOWN to be Hermitian-consistent (the spectrum of a real field):
        # true of everything forward() and our kernels produce, dropped when the caller gets the buffer.
        # The reference's x-space round trips (MHD RHS physics.py:797-815, max_square fields.py:153-157)
        # replace a spectrum by that of its real part; _hermitian_project() is that image.
        self._sym = True
        # vector components: True while the field this component belongs to is KNOWN to be solenoidal (the
        # fused RHS may then use the conservative products), False: known compressive (advective-form
        # policies), None: the caller has written the buffer since the last check (physics.verify_solenoidal)
        self._soln = True
        # The hand-out of the k-space buffer drops all of the above -- once.  A caller who KEEPS the tensor (the reference's
        # `uk = data['u']['x']['kspace']` is the live array) can write to it later; torch counts in-place writes to a tensor
        # and to every view of it (`_version`), so a buffer that has ever escaped is re-examined before a step whenever
        # that counter has moved since the last look (refresh_escaped; time_step._settle).  Writes torch cannot see
        # (DLPack / __cuda_array_interface__ consumers, raw pointers) need an explicit comp.touch().
        self._escaped = False
        self._seen_version = 0
        self._curr_space = "kspace"
        self.integrating_factor = None
        self.fwd_count = 0
        self.rev_count = 0
        self._sphere = None
        if dealiasing == "2/3 spherical":
            # '2/3 spherical' (representations.py:410-417): the library masks the bounding cube, this mask the rest
            self._sphere = (torch.sqrt(self.k2()) >= 2.0 / 3.0 * float(np.min(self.kny))).expand(self._k.shape)

    # ------------------------------------------------------------------ storage
    @property
    def xdata(self):
        if self._xdata is None:
            self._xdata = torch.zeros(tuple(int(n) for n in self.local_shape["xspace"]), dtype=torch.float64,
                                      device=self._plan.device)
        return self._xdata

    @property
    def kdata(self):
        """The k-space buffer.  Handing it out means the caller may modify it: the
        'zero outside the mask' knowledge is dropped (internal code uses _k)."""
        self._drop_knowledge()
        self._escaped = True
        self._seen_version = self._k._version
        return self._k

    def _drop_knowledge(self):
        self._clean = False
        self._checked = False
        self._sym = False
        self._soln = None

    def touch(self):
        """Tell the component that its k-space buffer has been modified behind torch's back (a write through a DLPack /
        CUDA-array-interface alias or a raw pointer): everything known about it is re-established before the next step."""
        self._drop_knowledge()

    def refresh_escaped(self):
        """A buffer that was handed out and has been written through torch since we last looked loses what we knew of it."""
        if self._escaped and self._k._version != self._seen_version:
            self._drop_knowledge()
            self._seen_version = self._k._version

    @property
    def data(self):
        return self.kdata if self._curr_space == "kspace" else self.xdata

    def __getitem__(self, space):
        self.require_space(space)
        return self.data

    def __setitem__(self, space, data):
        """Copy into the fixed buffer (its identity never changes, representations.py:144-166)."""
        if space == "xspace":
            target = self.xdata
        elif space == "kspace":
            target = self._k
            self._clean = isinstance(data, (float, complex, int)) and data == 0
            self._sym = self._clean
            self._soln = True if self._clean else None
            self._checked = False
        else:
            raise KeyError("space must be either xspace or kspace.")
        if isinstance(data, (float, complex, int)):
            target.fill_(data)
        else:
            if not torch.is_tensor(data):
                data = torch.as_tensor(np.asarray(data))
            if data.is_complex() and not target.is_complex():
                data = data.real
            if data.dim() == target.dim():
                data = data[tuple(slice(int(n)) for n in target.shape)]
            target.copy_(data, non_blocking=True)
        if space == "kspace":
            self._seen_version = self._k._version      # this write is accounted for (refresh_escaped)
        self._curr_space = space

    def require_space(self, space):
        if self._curr_space == space:
            return
        if space == "xspace":
            self.backward()
        elif space == "kspace":
            self.forward()
        else:
            raise ValueError("space must be either xspace or kspace.")

    # ------------------------------------------------------------------ host <-> device, retained modes only (extension)
    def _host_spectrum(self, host):
        if not (torch.is_tensor(host) and host.device.type == "cpu" and host.dtype == self._k.dtype and host.is_contiguous()
                and tuple(host.shape) == tuple(self._k.shape)):
            raise ValueError("host spectrum must be a contiguous CPU tensor of the local k-space shape and dtype.")
        return host

    def retained_parts(self):
        """Number of index boxes the retained modes of this component form (upload_retained / download_retained take a
        `part` in range(retained_parts()) to move one of them: finer-grained pipelining of host transfers)."""
        if self._sphere is not None or not self._static_k:
            return 1
        return len(self._plan.retained_boxes()[1])

    def upload_retained(self, host, part=None):
        """comp['kspace'] = host for a spectrum that is ZERO outside the dealias mask (the caller's promise, e.g. a state this
        package handed out): only the retained box crosses PCIe (30 % of the array under the 2/3 rule; include/ddl.h
        ddl_copy_boxes), the device buffer is zero elsewhere, and the component stays known-dealiased, so no check pass
        follows.  Asynchronous on the current stream for pinned memory.  Not in the reference (its arrays never leave the host);
        the assignment it replaces is representations.py:144-166."""
        host = self._host_spectrum(host)
        if self._sphere is not None or not self._static_k:
            self["kspace"] = host                      # masks that are not a box: the full copy
            return
        pl = self._plan
        if self._curr_space != "kspace" or not self._clean:
            check(lib.ddl_dealias(pl.handle, self._k.data_ptr(), _plan.current_stream()))      # zero outside the mask, once
        shape3, boxes = pl.retained_boxes()
        if part is not None:
            boxes = boxes[int(part):int(part) + 1]          # one box of the set (the caller uploads every part)
        check(lib.ddl_copy_boxes(self._k.data_ptr(), host.data_ptr(), shape3.ctypes.data, len(boxes), boxes.ctypes.data,
                                 self._k.element_size(), 1, _plan.current_stream()))
        self._curr_space = "kspace"
        self._clean = True
        self._checked = False
        self._sym = False
        self._soln = None

    def download_retained(self, host, part=None):
        """host <- comp['kspace'], retained box only; entries of `host` outside the mask are left as they are (zero them once).
        Raises ValueError when the spectrum carries content outside the mask (hydro states may): use ['kspace']."""
        host = self._host_spectrum(host)
        self.require_space("kspace")
        if self._sphere is not None or not self._static_k:
            host.copy_(self._k, non_blocking=True)
            return host
        if not self.verify_clean():
            raise ValueError("spectrum has content outside the dealias mask; download the full array (comp['kspace']).")
        shape3, boxes = self._plan.retained_boxes()
        if part is not None:
            boxes = boxes[int(part):int(part) + 1]
        check(lib.ddl_copy_boxes(host.data_ptr(), self._k.data_ptr(), shape3.ctypes.data, len(boxes), boxes.ctypes.data,
                                 self._k.element_size(), 0, _plan.current_stream()))
        return host

    def retained_bytes(self):
        """Bytes one upload_retained / download_retained moves."""
        _, boxes = self._plan.retained_boxes()
        return int(sum((b[1] - b[0]) * (b[3] - b[2]) * (b[5] - b[4]) for b in boxes)) * self._k.element_size()

    # ------------------------------------------------------------------ transforms
    @timer
    def forward(self):
        """x -> k: FFT / N_total, transposed out, then dealias (representations.py:335-345)."""
        if self._curr_space == "kspace":
            raise ValueError("Forward transform cannot be called from kspace.")
        pl = self._plan
        if pl.nranks > 1:
            pl.pipeline.forward(self.xdata, self._k)
        else:
            w = pl.transform_workspace()
            check(lib.ddl_forward(pl.handle, self.xdata.data_ptr(), self._k.data_ptr(), w.data_ptr(), w.numel(),
                                  _plan.current_stream()))
        if self._sphere is not None:
            self._k.masked_fill_(self._sphere, 0.0)
        self._curr_space = "kspace"
        self._clean = True
        self._sym = True
        self._soln = None           # an x-space field the caller wrote: divergence unknown
        self.fwd_count += 1

    @timer
    def backward(self):
        """k -> x: dealias the spectrum in place, then unnormalised inverse (:347-357)."""
        if self._curr_space == "xspace":
            raise ValueError("Backward transform cannot be called from xspace.")
        pl = self._plan
        if self._sphere is not None:
            self._k.masked_fill_(self._sphere, 0.0)
        if pl.nranks > 1:
            pl.pipeline.backward(self._k, self.xdata)
        else:
            w = pl.transform_workspace()
            check(lib.ddl_backward(pl.handle, self._k.data_ptr(), self.xdata.data_ptr(), w.data_ptr(), w.numel(),
                                   _plan.current_stream()))
        self._curr_space = "xspace"
        self._clean = True
        self.rev_count += 1

    fft = forward
    ifft = backward

    # ------------------------------------------------------------------ dealiasing
    def set_dealiasing(self, dealiasing):
        """Same option strings as representations.py:359-382; '2/3' and '2/3 cython' are one kernel."""
        n = self.global_shape["xspace"]
        if dealiasing in ("2/3", "2/3 cython"):
            self.nmodes = np.prod(2 * np.ceil(n / 3.0 - 1) + 1)
        elif dealiasing == "2/3 spherical":
            self.nmodes = None
        elif dealiasing in ("None", None, 0):
            self.nmodes = np.prod(2 * np.ceil(n / 2.0 - 1) + 1)
        else:
            raise NotImplementedError("Specified dealiasing method not implemented.")
        self._dealiasing = dealiasing

    def dealias(self):
        """Zero the modes outside the plan's mask, in place (dealias_cy_{2,3}d.pyx)."""
        self.require_space("kspace")
        check(lib.ddl_dealias(self._plan.handle, self._k.data_ptr(), _plan.current_stream()))
        if self._sphere is not None:
            self._k.masked_fill_(self._sphere, 0.0)
            self._seen_version = self._k._version      # our own write: accounted for (refresh_escaped)
        self._clean = True

    dealias_23_spherical = dealias
    dealias_23 = dealias
    dealias_23_cython = dealias

    def _hermitian_project(self):
        """k <- spectrum of the REAL PART of the field k describes: the kx = 0 plane becomes
        (k(ky,kz,0) + conj k(-ky,-kz,0)) / 2, every other stored plane is unchanged.  This is what a
        backward() + forward() round trip of the reference does to a spectrum whose kx = 0 plane is not
        Hermitian-consistent (c2r reads only the real part of the kx = 0 line values) -- e.g. the 3-D
        fields of turb_new, which never symmetrises that plane (init_cond.py:334-341).  Off the hot path:
        runs once for a buffer the caller has written, as tensor operations on one plane."""
        if self._sym:
            return
        self.require_space("kspace")
        d = self._k
        if self.ndim == 2:
            row = d[0]
            d[0] = 0.5 * (row + row.flip(0).roll(1, 0).conj())
            self._sym = True
            self._seen_version = d._version            # our own write: accounted for (refresh_escaped)
            return
        nranks = self._plan.nranks
        if nranks > 1:
            import torch.distributed as dist
            mine = torch.view_as_real(d[:, :, 0].contiguous())
            parts = [torch.empty_like(mine) for _ in range(nranks)]
            dist.all_gather(parts, mine)
            plane = torch.view_as_complex(torch.cat(parts, 0))
            if self._plan.ky_layout == "cyclic":      # rank-major rows -> global ky order
                ny = plane.shape[0]
                plane = plane.reshape(nranks, ny // nranks, -1).transpose(0, 1).reshape(ny, -1).contiguous()
        else:
            plane = d[:, :, 0]
        plane = 0.5 * (plane + plane.flip(0, 1).roll((1, 1), (0, 1)).conj())
        if nranks > 1:
            plane = plane[torch.as_tensor(self.local_rows["kspace"], device=d.device)]
        d[:, :, 0] = plane
        self._sym = True
        self._seen_version = d._version                # our own write: accounted for (refresh_escaped)

    def verify_clean(self):
        """Re-establish the 'zero outside the dealias mask' knowledge after the buffer was handed
        to the caller: one device pass over the masked-out slabs and one host read.  The
        integrators call it before a step so that states that ARE dealiased (every state that came
        out of forward(), div_free(), our own kernels) get the retained-modes-only sweeps and the
        fused stage kernel; states with genuine content outside the mask (hydro never dealiases
        its state) keep the full sweeps.  The verdict is cached until the next hand-out."""
        if self._clean or self._checked or self._curr_space != "kspace":
            return self._clean
        verify_clean_many([self])
        return self._clean

    def zero_nyquist(self):
        """Zero the Nyquist planes (representations.py:442-455)."""
        self.require_space("kspace")
        mask = None
        for name, kv in self.k.items():
            m = kv.abs() == float(self.kny[self.ktrans[name]])
            mask = m if mask is None else (mask | m)
        self.kdata.masked_fill_(mask.expand_as(self.kdata), 0.0)

    # ------------------------------------------------------------------ spectral helpers
    def deriv(self, dim):
        """i k_dim * data (representations.py:419-425); returns a fresh tensor."""
        self.require_space("kspace")
        out = torch.empty_like(self._k)
        check(lib.ddl_deriv(self._plan.handle, self._k.data_ptr(), out.data_ptr(), {"x": 0, "y": 1, "z": 2}[dim],
                            _plan.current_stream()))
        return out

    def k2(self, no_zero=False, set_zero=1.0):
        """|k|^2, summed in the reference's order (representations.py:427-440)."""
        k2 = torch.zeros(tuple(int(n) for n in self.local_shape["kspace"]), dtype=torch.float64, device=self._k.device)
        for kv in self.k.values():
            k2 += kv ** 2
        if no_zero:
            k2[k2 == 0] = set_zero
        return k2

    def find_mode(self, mode, exact=False):
        """LOCAL index of the mode closest to the physical wavevector `mode`, given in k-space
        axis order ((ky,kz,kx) / (kx,ky)); None if absent or owned by another rank
        (representations.py:247-288)."""
        idx = []
        for i in range(self.ndim):
            kv = self._plan.k_np[self.ktrans[i]]
            if i == 0:
                kv = kv[self.local_rows["kspace"]]
            if exact:
                hit = np.nonzero(kv == mode[i])[0]
            else:
                half = self.dk[i] / 2.0
                hit = np.nonzero((kv <= mode[i] + half) & (kv > mode[i] - half))[0]
            if len(hit) == 0:
                return None
            if len(hit) > 1:
                raise ValueError("Multiple modes tested true. This shouldn't happen.")
            idx.append(int(hit[0]))
        return tuple(idx)

    def enforce_hermitian(self):
        """Zero the Nyquist planes and overwrite the redundant half of the kx = 0 plane with
        the conjugate of its Hermitian partner (representations.py:457-503; with several ranks
        the plane is gathered, fixed and scattered back as the reference does, :480,:500)."""
        self.require_space("kspace")
        self.zero_nyquist()
        d = self.kdata
        if self.ndim == 2:
            ny = d.shape[1] // 2
            d[0, 0] = d[0, 0].real
            d[0, -ny:] = d[0, 1:ny + 1].flip(0).conj()
            return
        nranks = self._plan.nranks
        if nranks > 1:
            import torch.distributed as dist
            mine = torch.view_as_real(d[:, :, 0].contiguous())
            parts = [torch.empty_like(mine) for _ in range(nranks)]
            dist.all_gather(parts, mine)
            plane = torch.view_as_complex(torch.cat(parts, 0))
            if self._plan.ky_layout == "cyclic":      # rank-major rows -> global ky order
                ny = plane.shape[0]
                plane = plane.reshape(nranks, ny // nranks, -1).transpose(0, 1).reshape(ny, -1).contiguous()
        else:
            plane = d[:, :, 0]
        nyy, nyz = plane.shape[0] // 2, plane.shape[1] // 2
        plane[0, -nyz:] = plane[0, 1:nyz + 1].flip(0).conj()
        plane[-nyy:, 0] = plane[1:nyy + 1, 0].flip(0).conj()
        plane[-nyy:, 1:] = plane[1:nyy + 1, 1:].flip(0, 1).conj()
        if nranks > 1:
            d[:, :, 0] = plane[torch.as_tensor(self.local_rows["kspace"], device=d.device)]

    def zero_under_eps(self):
        self.require_space("kspace")
        self.kdata[self.kdata.abs() < self._eps["kspace"]] = 0.0

    # ------------------------------------------------------------------ grids / io
    def dx(self):
        return self.length / self.global_shape["xspace"]

    def xspace_grid(self, open=False):
        """Coordinates of the local x-space points, [ndim, ...] (representations.py:528-548)."""
        n = [int(v) for v in self.local_shape["xspace"]]
        dx = self.dx()
        axes = [torch.arange(n[i], dtype=torch.float64, device=self._k.device) * float(dx[i]) for i in range(self.ndim)]
        axes[0] = axes[0] + self.offset["xspace"] * float(dx[0])
        if open:
            return [a.reshape([-1 if j == i else 1 for j in range(self.ndim)]) for i, a in enumerate(axes)]
        return torch.stack(torch.meshgrid(*axes, indexing="ij"))

    def save(self, dataset):
        """Write the current data into an h5py-like dataset (representations.py:511-526)."""
        dataset[:] = self.data.cpu().numpy()
        dataset.attrs["space"] = self._curr_space


def verify_clean_many(comps):
    """verify_clean for several components of one plan with ONE device pass per eight of them and one host read
    (include/ddl.h: ddl_reduce_outside_mask)."""
    todo = [c for c in comps if not c._clean and not c._checked and c._curr_space == "kspace" and c._static_k]
    for i in range(0, len(todo), 8):
        group = todo[i:i + 8]
        pl = group[0]._plan
        out = torch.zeros(8, dtype=torch.float64, device=group[0]._k.device)
        arr = (C.c_void_p * len(group))(*[c._k.data_ptr() for c in group])
        check(lib.ddl_reduce_outside_mask(pl.handle, len(group), arr, out.data_ptr(), _plan.current_stream()))
        # deliberately rank-local (no collective): ranks may disagree (one of them printed a mode, say) and then take
        # different local tails (mask / full sweep vs fused sweep); the exchange sequence of the RHS pipeline is the same
        # on every path, so that is safe - a collective here could deadlock
        counts = out.cpu().tolist()
        for c, n in zip(group, counts):
            c._clean = not (n > 0 or n != n)
            c._checked = True
            if c._sphere is not None and c._clean:
                c._clean = not bool((c._k[c._sphere] != 0).any().item())      # '2/3 spherical': the bounding cube is not enough


class FourierShearRepresentation(FourierRepresentation):
    """Fourier representation in a shearing-box domain (representations.py:558-740).

    The wavenumber ky of every mode drifts with time, ky(kx, t) = ky0 - S kx t, wrapped once past the Nyquist value
    (:627-642), so `k['y']` is a dense array over (kx, ky) / (ky, 1, kx) and the 2/3 mask along y depends on kx and t.
    Transforms: x pass, the phase factor exp(+-i S kx y t) on the half-transformed lines, then the y (z) passes
    (fwd_np / rev_np :700-740) -- here the factor is applied INSIDE the x pass of ddl_forward / ddl_backward
    (include/ddl.h: ddl_set_shear), on a plan that keeps every ky row.  A compatibility path: the physics classes
    evaluate their right-hand side through the reference's unfused helpers for this representation and the
    integrators update with the array-factor stage kernel; one GPU or a slab decomposition (collective exchange)."""

    _static_k = False

    def __init__(self, sd, shape, length):
        FourierRepresentation.__init__(self, sd, shape, length)
        self.k = dict(self._plan.k)                       # private copy: k['y'] is this component's own, time dependent
        self._ky = self.k["y"].clone()
        self.k["y"] = (self._ky * torch.ones_like(self.k["x"])).contiguous()
        S = float(self.sd.parameters["shear_rate"])
        self._shear_rate = S
        self._wave_rate = S * self.k["x"]
        self._dy = float(self.dx()[self.xtrans["y"]])
        self._kshape64 = np.ascontiguousarray(self.local_shape["kspace"], dtype=np.int64)
        self._kny64 = np.ascontiguousarray(self.kny, dtype=np.float64)
        self._update_k()

    def _update_k(self):
        """Evolve the wavenumbers with the shear (representations.py:627-642): shift, wrap once, dealias."""
        ky = self.k["y"]
        torch.sub(self._ky, self._wave_rate * float(self.sd.time), out=ky)
        kny_y = float(self.kny[3 - self.ndim])
        ky.copy_(torch.where(ky <= -kny_y, ky + 2 * kny_y, ky))
        ky.copy_(torch.where(ky > kny_y, ky - 2 * kny_y, ky))
        self.dealias()

    def _set_shear(self):
        check(lib.ddl_set_shear(self._plan.handle, 1, self._shear_rate, float(self.sd.time), self._dy))

    def dealias(self):
        """The configured rule with the sheared ky (dealias_cy_2d.pyx:36-41, dealias_cy_3d.pyx:38-46;
        zero_nyquist :442-455 for FFT.dealiasing = None)."""
        self.require_space("kspace")
        if self._dealiasing in ("2/3", "2/3 cython"):
            # one launch of the kernel with the Cython kernels' own signature, dense-ky branch (include/ddl.h: ddl_dealias_array)
            kz = self.k["z"] if self.ndim == 3 else None
            check(lib.ddl_dealias_array(self.ndim, self._kshape64.ctypes.data_as(C.c_void_p), self._k.data_ptr(), self.k["x"].data_ptr(),
                                        self.k["y"].data_ptr(), kz.data_ptr() if kz is not None else None, 1,
                                        self._kny64.ctypes.data_as(C.c_void_p), _plan.current_stream()))
        else:
            mask = None
            for name, kv in self.k.items():
                m = kv.abs() == float(self.kny[self.ktrans[name]])
                mask = m if mask is None else (mask | m)
            self._k.masked_fill_(mask.expand_as(self._k), 0.0)
        self._clean = True

    dealias_23 = dealias
    dealias_23_cython = dealias

    def zero_nyquist(self):
        self.require_space("kspace")
        mask = None
        for name, kv in self.k.items():
            m = kv.abs() == float(self.kny[self.ktrans[name]])
            mask = m if mask is None else (mask | m)
        self.kdata.masked_fill_(mask.expand_as(self._k), 0.0)

    @timer
    def forward(self):
        if self._curr_space == "kspace":
            raise ValueError("Forward transform cannot be called from kspace.")
        self._set_shear()
        FourierRepresentation.forward(self)
        self.dealias()

    @timer
    def backward(self):
        if self._curr_space == "xspace":
            raise ValueError("Backward transform cannot be called from xspace.")
        self.dealias()
        self._set_shear()
        FourierRepresentation.backward(self)

    fft = forward
    ifft = backward

    def deriv(self, dim):
        if dim != "y":
            return FourierRepresentation.deriv(self, dim)
        self.require_space("kspace")
        return self._k * (1j * self.k["y"])

    def verify_clean(self):
        return self._clean

    def _hermitian_project(self):
        raise NotImplementedError("FourierShearRepresentation: not on the fused path")


class ChebyshevRepresentation(FourierRepresentation):
    """Placeholder of the reference (representations.py:743-744: `class ChebyshevRepresentation(FourierRepresentation): pass`)."""
