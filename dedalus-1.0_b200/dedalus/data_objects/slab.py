"""Slab-decomposed 3-D transforms and fused RHS of one rank (one process per GPU).

Partition exactly as the reference / FFTW-MPI (dedalus/utils/fftw/_fftw.pyx:114-148,
representations.py:180-186,231-233): x-space split along z, k-space along ky.  FFTW-MPI hides
the global transpose inside every ``fftw_execute`` (_fftw.pyx:232-234, 272-304: one blocking
MPI all-to-all per transform); here the pipeline is cut into the library's phases
(include/ddl.h: ddl_slab_*) and the exchange between the z and y passes is an asynchronous
``all_to_all_single`` of contiguous per-peer blocks, issued field by field so that field f's
exchange overlaps field f+1's z / y passes (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  The z-pass kernels already write / read the peer-blocked layout, so no pack or unpack
pass exists on either side of the exchange.

The class is device-agnostic on purpose: `lib` is the C-ABI library (libddl_b200.so in the
product; the tests hand in the host-emulation build and CPU tensors).
"""
import ctypes as C

import torch
import torch.distributed as dist

_COUNTS = {0: (3, 6), 1: (4, 9), 2: (6, 9)}     # physics id -> (inverse, forward) transforms, 3-D


def _ptrs(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _Done(object):
    def wait(self):
        return True


class SlabPipeline(object):
    def __init__(self, lib, handle, device, group=None, stream=None):
        self.lib, self.h, self.device, self.group = lib, handle, device, group
        self._stream = stream or (lambda: None)
        info = (C.c_int64 * 16)()
        self._check(lib.ddl_slab_info(handle, info))
        (self.P, self.rank, self.nzl, self.nyl, self.cyl, self.cy0, self.cy, self.cz, self.cx, self.nkx,
         self.n_ks, self.n_xs, self.n_b, self.n_e, self.z0, self.ky0) = [int(v) for v in info]
        rows = (C.c_int64 * self.P)()
        self._check(lib.ddl_slab_rows(handle, rows))
        self.rows = [int(v) for v in rows]
        blk = self.nzl * self.cx
        # elements per field: what I send to every peer / what peer r sends me (inverse direction)
        self.to_peer = [self.cyl * blk] * self.P
        self.from_peer = [r * blk for r in self.rows]
        self._bufs = {}
        self.exchanges = 0

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.ddl_last_error().decode())

    # ------------------------------------------------------------------ buffers
    def buffers(self, ni, no):
        key = (ni, no)
        b = self._bufs.get(key)
        if b is None:
            nmax = max(ni, no)
            z = lambda n: torch.zeros(max(int(n), 1), dtype=torch.complex128, device=self.device)
            ks = z(nmax * self.n_ks)
            xs = ks if self.P == 1 else z(nmax * self.n_xs)     # one rank: the two layouts coincide
            be = z(max(ni * self.n_b, no * self.n_e))
            c = z(no * self.n_b)
            cut = lambda t, n, k: [t[i * n:(i + 1) * n] for i in range(k)]
            b = {"ks": cut(ks, self.n_ks, nmax), "xs": cut(xs, self.n_xs, nmax), "b": cut(be, self.n_b, ni),
                 "e": cut(be, self.n_e, no), "c": cut(c, self.n_b, no)}
            self._bufs = {key: b}          # keep one set alive (the largest use in practice)
        return b

    # ------------------------------------------------------------------ exchange
    def _exchange(self, dst, src, inverse):
        """inverse: k-side (peer-blocked) -> x-side; forward: x-side -> k-side.  Asynchronous."""
        if self.P == 1:
            return _Done()
        self.exchanges += 1
        out_split, in_split = (self.from_peer, self.to_peer) if inverse else (self.to_peer, self.from_peer)
        return dist.all_to_all_single(torch.view_as_real(dst), torch.view_as_real(src), out_split, in_split,
                                      group=self.group, async_op=True)

    # ------------------------------------------------------------------ transforms
    def backward(self, k, x):
        """k (local slab, dealiased in place) -> x (local planes), unnormalised (representations.py:347-357)."""
        lib, h, st = self.lib, self.h, self._stream()
        b = self.buffers(1, 1)
        self._check(lib.ddl_dealias(h, k.data_ptr(), st))
        self._check(lib.ddl_slab_zinv(h, 1, _ptrs([k]), _ptrs([b["ks"][0]]), st))
        self._exchange(b["xs"][0], b["ks"][0], True).wait()
        self._check(lib.ddl_slab_yinv(h, 1, _ptrs([b["xs"][0]]), _ptrs([b["b"][0]]), st))
        self._check(lib.ddl_slab_xc2r(h, b["b"][0].data_ptr(), x.data_ptr(), st))

    def forward(self, x, k):
        """x (local planes) -> k (local slab): normalised, transposed, dealiased (representations.py:335-345)."""
        lib, h, st = self.lib, self.h, self._stream()
        b = self.buffers(1, 1)
        self._check(lib.ddl_slab_xr2c(h, x.data_ptr(), b["c"][0].data_ptr(), st))
        self._check(lib.ddl_slab_yfwd(h, 1, _ptrs([b["c"][0]]), _ptrs([b["xs"][0]]), st))
        self._exchange(b["ks"][0], b["xs"][0], False).wait()
        self._check(lib.ddl_slab_zfwd(h, 1, _ptrs([b["ks"][0]]), _ptrs([k]), 1, st))
        self._check(lib.ddl_dealias(h, k.data_ptr(), st))

    # ------------------------------------------------------------------ fused RHS
    def rhs(self, physics_id, params, state, deriv, dealias_state, zero_fill):
        """deriv = RHS(state) (physics.py:527-599 / 664-712 / 770-819), local slabs in and out."""
        lib, h, st = self.lib, self.h, self._stream()
        ni, no = _COUNTS[physics_id]
        b = self.buffers(ni, no)
        pp = C.byref(params)
        if dealias_state:
            for t in state:
                self._check(lib.ddl_dealias(h, t.data_ptr(), st))
        if zero_fill:
            for t in deriv:
                self._check(lib.ddl_dealias(h, t.data_ptr(), st))
        ks, xs = b["ks"], b["xs"]
        # inverse: z pass of field f, its exchange in flight while field f+1 is transformed
        pending = []
        for f in range(ni):
            self._check(lib.ddl_slab_zinv(h, 1, _ptrs([state[f]]), _ptrs([ks[f]]), st))
            pending.append(self._exchange(xs[f], ks[f], True))
        for f in range(ni):
            pending[f].wait()
            self._check(lib.ddl_slab_yinv(h, 1, _ptrs([xs[f]]), _ptrs([b["b"][f]]), st))
        self._check(lib.ddl_slab_xfused(h, physics_id, pp, _ptrs(b["b"][:ni]), _ptrs(b["c"][:no]), st))
        # forward: y pass of product f, exchange, z pass
        pending = []
        for f in range(no):
            self._check(lib.ddl_slab_yfwd(h, 1, _ptrs([b["c"][f]]), _ptrs([xs[f]]), st))
            pending.append(self._exchange(ks[f], xs[f], False))
        for f in range(no):
            pending[f].wait()
            self._check(lib.ddl_slab_zfwd(h, 1, _ptrs([ks[f]]), _ptrs([b["e"][f]]), 0, st))
        self._check(lib.ddl_slab_assemble(h, physics_id, pp, _ptrs(b["e"][:no]), _ptrs(state), _ptrs(deriv), st))
