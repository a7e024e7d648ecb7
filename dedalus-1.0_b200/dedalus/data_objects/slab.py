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

# physics id -> (inverse, forward) transforms, 3-D; 3..5: advective-form policies for non-solenoidal states
# (include/ddl.h DDL_*_ADV: the divergence spectra are extra inverse inputs, their products extra forward outputs)
_COUNTS = {0: (3, 6), 1: (4, 9), 2: (6, 9), 3: (4, 9), 4: (5, 13), 5: (8, 12)}


def _ptrs(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def arena_layout(P, me, rows, nzl, cx, nz, nmax, el=16):
    """Byte offsets of the slab exchange inside the ranks' arenas (pure arithmetic; tests/test_slab_layout.py).

    Every arena's data part is [nmax x-side fields | nmax k-side fields]:
      x-side field f of any rank : [cy][nzl][cx], rows in owner-major order  (cy = sum(rows))
      k-side field f of rank r   : [peer][rows[r]][nzl][cx]
    Returns a dict with, for this rank `me`:
      xs[f], ks[f]            offsets of its own fields
      inv[f] = (dst_rank, src_off, dst_off, nbytes) lists: my k-side block s -> rows of `me` in rank s's x-side field f
      fwd[f] = the mirror: rows of rank s in my x-side field f -> block `me` of rank s's k-side field f
      zinv_peer[f][s], yfwd_peer[f][s]   block bases (relative to rank s's data base) for the peer-store passes:
            z pass: row z of my local compact row j goes to  base + (j*nzl + z % nzl)*cx
            y pass: x-side position p of plane zl goes to     base + (p*nzl + zl)*cx      (base absorbs -cy0[s])
    """
    blk = nzl * cx
    cy = sum(rows)
    cy0 = [sum(rows[:r]) for r in range(P)]
    n_xs = cy * blk
    n_ks = [r * nz * cx for r in rows]
    xs_total = nmax * n_xs * el
    order = [(me + i) % P for i in range(P)]            # start with myself, then staggered peers
    out = {"xs": [f * n_xs * el for f in range(nmax)], "ks": [xs_total + f * n_ks[me] * el for f in range(nmax)],
           "bytes": xs_total + nmax * n_ks[me] * el, "order": order, "inv": [], "fwd": [], "zinv_peer": [], "yfwd_peer": []}
    for f in range(nmax):
        out["inv"].append((order,
                           [xs_total + (f * n_ks[me] + s * rows[me] * blk) * el for s in order],
                           [(f * n_xs + cy0[me] * blk) * el for s in order],
                           [rows[me] * blk * el for s in order]))
        out["fwd"].append((order,
                           [(f * n_xs + cy0[s] * blk) * el for s in order],
                           [xs_total + (f * n_ks[s] + me * rows[s] * blk) * el for s in order],
                           [rows[s] * blk * el for s in order]))
        out["zinv_peer"].append([(f * n_xs + cy0[me] * blk) * el for s in range(P)])
        out["yfwd_peer"].append([xs_total + (f * n_ks[s] + (me * rows[s] - cy0[s]) * blk) * el for s in range(P)])
    return out


class _Done(object):
    def wait(self):
        return True


class _Raw(object):
    """A device address inside the IPC arena, quacking like a tensor for the C calls."""

    def __init__(self, ptr):
        self.ptr = int(ptr)

    def data_ptr(self):
        return self.ptr


class _Ticket(object):
    def __init__(self, pipe, seq):
        self.pipe, self.seq = pipe, seq

    def wait(self):
        p = self.pipe
        p._check(p.lib.ddl_p2p_wait(p._p2p, self.seq, p._stream()))
        return True


class SlabPipeline(object):
    def __init__(self, lib, handle, device, group=None, stream=None, exchange="collective"):
        """exchange of the RHS pipeline (the plain transforms always use the collective):
        "collective" = torch.distributed all_to_all_single (NCCL / gloo);
        "p2p"  = CUDA-IPC arenas + copy-engine pushes + arrival flags (csrc/p2p.cu);
        "peer" = the producing z / y pass stores straight into the peers' arenas over NVLink
                 (no send buffer, no copy) + arrival flags;
        "push" = the passes write locally and a small SM kernel on a second stream pushes the blocks to the peers
                 (csrc/p2p.cu p2p_push_kernel) while the next passes run: the all-to-all overlapped with the 1-D transforms."""
        self.lib, self.h, self.device, self.group = lib, handle, device, group
        self.exchange_kind = exchange
        self._p2p = None
        self._stream = stream or (lambda: None)
        info = (C.c_int64 * 32)()
        self._check(lib.ddl_slab_info(handle, info))
        (self.P, self.rank, self.nzl, self.nyl, self.cyl, self.cy0, self.cy, self.cz, self.cx, self.nkx,
         self.n_ks, self.n_xs, self.n_b, self.n_e, self.z0, self.ky0, self.ky_layout) = [int(v) for v in info[:17]]
        rows = (C.c_int64 * self.P)()
        self._check(lib.ddl_slab_rows(handle, rows))
        self.rows = [int(v) for v in rows]
        blk = self.nzl * self.cx
        # elements per field: what I send to every peer / what peer r sends me (inverse direction)
        self.to_peer = [self.cyl * blk] * self.P
        self.from_peer = [r * blk for r in self.rows]
        self._bufs = {}
        self.exchanges = 0
        import os
        self.chunks = int(os.environ.get("DEDALUS_SLAB_CHUNKS", "2"))     # plane chunks of the forward x / y passes (peer exchange)
        # inverse half: "batched" = one z pass and one y pass for all fields; "groups:G" = G field groups, the y pass of
        # group g on the side stream while the (NVLink-bound) z pass of group g+1 pushes its rows; "fields" = one group per field
        inv = os.environ.get("DEDALUS_SLAB_INVERSE", "batched")
        self.inverse_batched = inv == "batched"
        self.inverse_groups = int(inv.split(":")[1]) if inv.startswith("groups:") else 0
        # CTAs the peer-store passes may occupy (include/ddl.h "peer_pass_ctas"; 0 = one per tile) and whether the stream
        # they run on outranks the main stream, so that their few CTAs are placed as soon as an SM has room
        self.peer_ctas = int(os.environ.get("DEDALUS_PEER_CTAS", "0"))
        self.push_ctas = int(os.environ.get("DEDALUS_PUSH_CTAS", "32"))       # CTAs of the push kernel ("push" exchange)
        self.side_priority = os.environ.get("DEDALUS_SIDE_PRIORITY", "0") == "1"
        # tests only (tests/test_gpu_slab.py skew cases): delay this rank's GPU work by rank- and call-dependent amounts at the phase
        # boundaries of the RHS, so that the ranks drift against each other by whole passes -- the arrival-flag protocol has no
        # "buffer free" credits and must be correct under any such drift
        self.test_skew = int(os.environ.get("DEDALUS_TEST_SKEW", "0"))
        self._skew_calls = 0
        self.trace = None               # profiling only: list collecting (label, torch.cuda.Event) marks of rhs()
        self.skip_exchange = False      # profiling only (profiles/slab_breakdown.py): time the passes without the all-to-all

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.ddl_last_error().decode())

    def _mark(self, label):
        if self.test_skew:
            # a different rank is the slow one at every mark, by up to test_skew milliseconds (torch.cuda._sleep: GPU-side spin)
            self._skew_calls += 1
            k = (self._skew_calls * 7 + self.rank * 3) % (self.P + 1)
            if k:
                torch.cuda._sleep(int(k * self.test_skew * 1.5e6 / self.P))
        if self.trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.trace.append((label, ev))

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
            self._bufs[key] = b
        return b

    # ------------------------------------------------------------------ p2p arena (RHS pipeline)
    def _p2p_setup(self, nmax):
        """Arena data layout (after the flag header, csrc/p2p.cu), identical on every rank up to the
        peer's own k-side size: [nmax x-side fields | nmax k-side fields]."""
        lib, P, me = self.lib, self.P, self.rank
        lay = arena_layout(P, me, self.rows, self.nzl, self.cx, self.nzl * P, nmax)
        import os
        if os.environ.get("DEDALUS_P2P_TIMEOUT"):       # seconds a pass waits for a late peer before it traps; 0 = for ever
            self._check(lib.ddl_set_option(b"p2p_timeout_s", int(float(os.environ["DEDALUS_P2P_TIMEOUT"]))))
        ctx = C.c_void_p()
        handle = C.create_string_buffer(64)
        self._check(lib.ddl_p2p_create(C.byref(ctx), P, me, lay["bytes"], handle))
        gathered = [None] * P
        dist.all_gather_object(gathered, handle.raw, group=self.group)
        self._check(lib.ddl_p2p_connect(ctx, b"".join(gathered)))
        self._p2p = ctx
        self._p2p_nmax = nmax
        self._p2p_layout = lay
        base = int(lib.ddl_p2p_base(ctx))
        self._p2p_xs = [_Raw(base + o) for o in lay["xs"]]
        self._p2p_ks = [_Raw(base + o) for o in lay["ks"]]
        arr = lambda vals, t: (t * len(vals))(*vals)
        self._p2p_lists = {}
        for f in range(nmax):
            for inverse, key in ((True, "inv"), (False, "fwd")):
                ranks, src, dst, nb = lay[key][f]
                self._p2p_lists[(f, inverse)] = (arr(ranks, C.c_int), arr(src, C.c_int64), arr(dst, C.c_int64), arr(nb, C.c_int64))
        # tables of the peer-store passes (exchange fused into the z / y pass): [f][s] block bases
        pb = [int(lib.ddl_p2p_peer_base(ctx, s)) for s in range(P)]
        zt = [[pb[s] + lay["zinv_peer"][f][s] for s in range(P)] for f in range(nmax)]
        yt = [[pb[s] + lay["yfwd_peer"][f][s] for s in range(P)] for f in range(nmax)]
        self._zinv_tab = torch.tensor(zt, dtype=torch.int64, device=self.device)
        self._yfwd_tab = torch.tensor(yt, dtype=torch.int64, device=self.device)
        self._side = torch.cuda.Stream(device=self.device, priority=-1 if self.side_priority else 0)
        if self.peer_ctas:
            self._check(lib.ddl_set_option(b"peer_pass_ctas", self.peer_ctas))
        dist.barrier(group=self.group)

    def _signal(self):
        self.exchanges += 1
        seq = self.lib.ddl_p2p_signal(self._p2p, self._stream())
        if seq <= 0:
            self._check(int(seq) or -1)
        return _Ticket(self, seq)

    def _rhs_peer(self, physics_id, pp, state, deriv, ni, no, fuse=None):
        """RHS with the exchange fused into the producing passes: the inverse z pass and the
        forward y pass store their output rows straight into the owning rank's arena over
        NVLink; a second stream runs the consuming passes of field f while the producing pass
        of field f+1 (NVLink-bound) is in flight."""
        lib, h = self.lib, self.h
        b = self._p2p_buffers(ni, no)
        ks, xs = b["ks"], b["xs"]
        main = torch.cuda.current_stream()
        side = self._side
        el = 8 * self.P          # bytes per table row
        zt, yt = self._zinv_tab.data_ptr(), self._yfwd_tab.data_ptr()
        self._mark("start")
        side.wait_stream(main)
        if self.inverse_batched:
            # one launch per pass for all fields: at 8 ranks a single field is only ~3 waves of CTAs
            self._check(lib.ddl_slab_zinv_peer(h, ni, _ptrs(state[:ni]), zt, main.cuda_stream))
            t = self._signal()
            self._mark("z_inv")
            t.wait()
            self._check(lib.ddl_slab_yinv(h, ni, _ptrs(xs[:ni]), _ptrs(b["b"][:ni]), main.cuda_stream))
        else:
            ng = min(self.inverse_groups, ni) if self.inverse_groups else ni
            bounds = [(g * ni) // ng for g in range(ng + 1)]
            done = []
            for g in range(ng):
                f0, f1 = bounds[g], bounds[g + 1]
                self._check(lib.ddl_slab_zinv_peer(h, f1 - f0, _ptrs(state[f0:f1]), zt + f0 * el, main.cuda_stream))
                t = self._signal()
                ev = torch.cuda.Event()
                ev.record(main)
                done.append((t, ev))
            self._mark("z_inv")
            with torch.cuda.stream(side):
                for g in range(ng):
                    f0, f1 = bounds[g], bounds[g + 1]
                    t, ev = done[g]
                    side.wait_event(ev)             # my own block is written by my own pass
                    t.wait()                        # the peers' blocks: arrival flags
                    self._check(lib.ddl_slab_yinv(h, f1 - f0, _ptrs(xs[f0:f1]), _ptrs(b["b"][f0:f1]), side.cuda_stream))
            main.wait_stream(side)
        self._mark("wait+y_inv")
        # forward, chunked over the local planes: the x pass of chunk c+1 (compute-bound) runs on
        # the main stream while the y pass of chunk c pushes its rows to the peers (NVLink-bound)
        # on the side stream; one arrival signal once every chunk has been stored
        nch = self.chunks if self.nzl % self.chunks == 0 else 1
        zc = self.nzl // nch
        bin_, cout = _ptrs(b["b"][:ni]), _ptrs(b["c"][:no])
        for c in range(nch):
            self._check(lib.ddl_slab_xfused_planes(h, physics_id, pp, bin_, cout, c * zc, zc, main.cuda_stream))
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            self._check(lib.ddl_slab_yfwd_peer(h, no, cout, yt, c * zc, zc, side.cuda_stream))
        self._mark("x_fused")
        with torch.cuda.stream(side):
            t = self._signal()
            t.wait()
            self._check(lib.ddl_slab_zfwd(h, no, _ptrs(ks[:no]), _ptrs(b["e"][:no]), 0, side.cuda_stream))
        main.wait_stream(side)
        self._mark("y_fwd+z_fwd")
        self._assemble(physics_id, pp, b["e"][:no], state, deriv, fuse, main.cuda_stream)
        self._mark("assemble")

    # ------------------------------------------------------------------ "push" exchange
    def _push_table(self, key):
        """ctypes argument block of ddl_p2p_push for ("inv", f0, f1) -- my k-side blocks of fields f0..f1 to the peers' x-side
        rows -- or ("fwd", nf, z0, nzc) -- planes [z0, z0 + nzc) of the peers' rows of my x-side fields 0..nf to my block of
        their k-side fields.  Offsets from arena_layout (tests/test_slab_layout.py); built once per key."""
        tabs = self.__dict__.setdefault("_push_tabs", {})
        t = tabs.get(key)
        if t is None:
            lay = self._p2p_layout
            el, blk = 16, self.nzl * self.cx
            ranks, src, dst, rb, nr, pitch = [], [], [], [], [], []
            if key[0] == "inv":
                for f in range(key[1], key[2]):
                    order, so, do, nb = lay["inv"][f]
                    for i, s in enumerate(order):
                        ranks.append(s); src.append(so[i]); dst.append(do[i]); rb.append(nb[i]); nr.append(1); pitch.append(nb[i])
            else:
                _, nf, z0, nzc = key
                for f in range(nf):
                    order, so, do, nb = lay["fwd"][f]
                    for i, s in enumerate(order):
                        ranks.append(s); src.append(so[i] + z0 * self.cx * el); dst.append(do[i] + z0 * self.cx * el)
                        rb.append(nzc * self.cx * el); nr.append(self.rows[s]); pitch.append(blk * el)
            arr = lambda vals, t: (t * len(vals))(*vals)
            t = tabs[key] = (len(ranks), arr(ranks, C.c_int), arr(src, C.c_int64), arr(dst, C.c_int64), arr(rb, C.c_int64),
                             arr(nr, C.c_int64), arr(pitch, C.c_int64))
        return t

    def _push(self, key, publish, stream):
        n, ranks, src, dst, rb, nr, pitch = self._push_table(key)
        seq = self.lib.ddl_p2p_push(self._p2p, n, ranks, src, dst, rb, nr, pitch, self.push_ctas, 1 if publish else 0, stream.cuda_stream)
        if seq < 0:
            self._check(int(seq))
        if publish:
            self.exchanges += 1
        return seq

    def _rhs_push(self, physics_id, pp, state, deriv, ni, no, fuse=None):
        """RHS with the all-to-all carried by a small SM kernel on a second (high-priority) stream: every pass reads and writes
        local memory only (so it runs at its HBM / FP64 speed), the blocks it produced are pushed to the peers over NVLink
        while the following passes run, and a pass waits for arrival flags only.  Inverse half: field groups (z pass of
        group g+1 and y pass of group g-1 overlap the push of group g); forward half: plane chunks (the push of chunk c
        overlaps the x and y passes of chunk c+1)."""
        lib, h = self.lib, self.h
        b = self._p2p_buffers(ni, no)
        ks, xs = b["ks"], b["xs"]
        main, side = torch.cuda.current_stream(), self._side
        self._mark("start")
        ng = min(self.inverse_groups, ni) if self.inverse_groups else (1 if self.inverse_batched else ni)
        bounds = [(g * ni) // ng for g in range(ng + 1)]
        sent = []
        for g in range(ng):
            f0, f1 = bounds[g], bounds[g + 1]
            self._check(lib.ddl_slab_zinv(h, f1 - f0, _ptrs(state[f0:f1]), _ptrs(ks[f0:f1]), main.cuda_stream))
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            seq = self._push(("inv", f0, f1), True, side)
            ev = torch.cuda.Event()
            ev.record(side)
            sent.append((seq, ev))
        self._mark("z_inv")
        for g in range(ng):
            f0, f1 = bounds[g], bounds[g + 1]
            seq, ev = sent[g]
            main.wait_event(ev)                                          # my own block (a local copy inside the push)
            self._check(lib.ddl_p2p_wait(self._p2p, seq, main.cuda_stream))   # the peers' blocks: arrival flags
            self._check(lib.ddl_slab_yinv(h, f1 - f0, _ptrs(xs[f0:f1]), _ptrs(b["b"][f0:f1]), main.cuda_stream))
        self._mark("wait+y_inv")
        nch = self.chunks if self.nzl % self.chunks == 0 else 1
        zc = self.nzl // nch
        bin_, cout, xout = _ptrs(b["b"][:ni]), _ptrs(b["c"][:no]), _ptrs(xs[:no])
        for c in range(nch):
            self._check(lib.ddl_slab_xfused_planes(h, physics_id, pp, bin_, cout, c * zc, zc, main.cuda_stream))
            self._check(lib.ddl_slab_yfwd_planes(h, no, cout, xout, c * zc, zc, main.cuda_stream))
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            seq = self._push(("fwd", no, c * zc, zc), c == nch - 1, side)
        self._mark("x_fused")
        ev = torch.cuda.Event()
        ev.record(side)
        main.wait_event(ev)
        self._check(lib.ddl_p2p_wait(self._p2p, seq, main.cuda_stream))
        self._check(lib.ddl_slab_zfwd(h, no, _ptrs(ks[:no]), _ptrs(b["e"][:no]), 0, main.cuda_stream))
        self._mark("y_fwd+z_fwd")
        self._assemble(physics_id, pp, b["e"][:no], state, deriv, fuse, main.cuda_stream)
        self._mark("assemble")

    def _exchange_p2p(self, f, inverse):
        self.exchanges += 1
        ranks, src, dst, nb = self._p2p_lists[(f, inverse)]
        seq = self.lib.ddl_p2p_exchange(self._p2p, self.P, ranks, src, dst, nb, self._stream())
        if seq <= 0:
            self._check(int(seq) or -1)
        return _Ticket(self, seq)

    def _p2p_buffers(self, ni, no):
        """RHS buffer set whose exchange endpoints (k-side / x-side) live in the IPC arena."""
        nmax = max(ni, no)
        if self._p2p is None:
            self._p2p_setup(13)      # the largest field count of any policy: 13 products of the advective-form Boussinesq RHS
        if nmax > self._p2p_nmax:
            raise RuntimeError("p2p arena holds %d fields, %d requested" % (self._p2p_nmax, nmax))
        key = ("p2p", ni, no)
        b = self._bufs.get(key)
        if b is None:
            z = lambda n: torch.zeros(max(int(n), 1), dtype=torch.complex128, device=self.device)
            be = z(max(ni * self.n_b, no * self.n_e))
            c = z(no * self.n_b)
            cut = lambda t, n, k: [t[i * n:(i + 1) * n] for i in range(k)]
            b = {"ks": self._p2p_ks[:nmax], "xs": self._p2p_xs[:nmax], "b": cut(be, self.n_b, ni),
                 "e": cut(be, self.n_e, no), "c": cut(c, self.n_b, no)}
            self._bufs[key] = b
        return b

    # ------------------------------------------------------------------ exchange
    def _exchange(self, dst, src, inverse):
        """inverse: k-side (peer-blocked) -> x-side; forward: x-side -> k-side.  Asynchronous."""
        if self.P == 1 or self.skip_exchange:
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

    def max_square(self, physics_id, params, state, out2, dealias_state):
        """out2[0:2] = max u_i(x)^2, max B_i(x)^2 over the local planes: the inverse half of the RHS
        pipeline with the CFL capture of the x pass on (include/ddl.h: ddl_rhs_capture_max); the
        caller reduces over ranks."""
        lib, h, st = self.lib, self.h, self._stream()
        ni, no = _COUNTS[physics_id]
        b = self.buffers(ni, no)
        if dealias_state:
            for t in state:
                self._check(lib.ddl_dealias(h, t.data_ptr(), st))
        pending = []
        for f in range(ni):
            self._check(lib.ddl_slab_zinv(h, 1, _ptrs([state[f]]), _ptrs([b["ks"][f]]), st))
            pending.append(self._exchange(b["xs"][f], b["ks"][f], True))
        for f in range(ni):
            pending[f].wait()
            self._check(lib.ddl_slab_yinv(h, 1, _ptrs([b["xs"][f]]), _ptrs([b["b"][f]]), st))
        self._check(lib.ddl_rhs_capture_max(h, out2.data_ptr()))
        try:
            self._check(lib.ddl_slab_xfused(h, physics_id, C.byref(params), _ptrs(b["b"][:ni]), _ptrs(b["c"][:no]), st))
        finally:
            self._check(lib.ddl_rhs_capture_max(h, None))

    # ------------------------------------------------------------------ fused RHS
    def _assemble(self, physics_id, pp, e, state, deriv, fuse, st):
        lib, h = self.lib, self.h
        if fuse is not None:
            self._check(lib.ddl_slab_assemble_stage(h, physics_id, pp, _ptrs(e), _ptrs(state), C.byref(fuse), st))
        else:
            self._check(lib.ddl_slab_assemble(h, physics_id, pp, _ptrs(e), _ptrs(state), _ptrs(deriv), st))

    def rhs(self, physics_id, params, state, deriv, dealias_state, zero_fill, fuse=None, ncomp=None):
        """deriv = RHS(state) (physics.py:527-599 / 664-712 / 770-819), local slabs in and out.
        Advective-form policies (physics_id >= 3): `state` = the ncomp state slabs followed by the scratch
        slabs for the divergence spectra; with the "peer" exchange they go through the same fused peer-store
        pipeline as the solenoidal policies (more fields), with "p2p" / "push" through the collective."""
        lib, h, st = self.lib, self.h, self._stream()
        ni, no = _COUNTS[physics_id]
        adv = physics_id >= 3
        p2p = self.exchange_kind == "p2p" and self.P > 1 and not self.skip_exchange and not adv
        # the advective-form policies (states that are not solenoidal) take the peer-store exchange too: the arenas hold 13 fields
        peer = self.exchange_kind == "peer" and self.P > 1 and not self.skip_exchange
        push = self.exchange_kind == "push" and self.P > 1 and not self.skip_exchange and not adv
        pp = C.byref(params)
        if dealias_state:
            for t in state[:ncomp]:
                self._check(lib.ddl_dealias(h, t.data_ptr(), st))
        if adv:
            self._check(lib.ddl_slab_theta(h, physics_id, _ptrs(state), st))
        if zero_fill:
            for t in deriv:
                self._check(lib.ddl_dealias(h, t.data_ptr(), st))
        self.last_rhs_path = "peer" if peer else "push" if push else "p2p" if p2p else "collective"     # what carried the last exchange (tests)
        if peer:
            return self._rhs_peer(physics_id, pp, state, deriv, ni, no, fuse)
        if push:
            return self._rhs_push(physics_id, pp, state, deriv, ni, no, fuse)
        b = dict(self.buffers(ni, no)) if not p2p else dict(self._p2p_buffers(ni, no))
        ks, xs = b["ks"], b["xs"]
        self._mark("start")
        # inverse: z pass of field f, its exchange in flight while field f+1 is transformed
        pending = []
        for f in range(ni):
            self._check(lib.ddl_slab_zinv(h, 1, _ptrs([state[f]]), _ptrs([ks[f]]), st))
            pending.append(self._exchange_p2p(f, True) if p2p else self._exchange(xs[f], ks[f], True))
        self._mark("z_inv")
        for f in range(ni):
            pending[f].wait()
            self._check(lib.ddl_slab_yinv(h, 1, _ptrs([xs[f]]), _ptrs([b["b"][f]]), st))
        self._mark("wait+y_inv")
        self._check(lib.ddl_slab_xfused(h, physics_id, pp, _ptrs(b["b"][:ni]), _ptrs(b["c"][:no]), st))
        self._mark("x_fused")
        # forward: y pass of product f, exchange, z pass
        pending = []
        for f in range(no):
            self._check(lib.ddl_slab_yfwd(h, 1, _ptrs([b["c"][f]]), _ptrs([xs[f]]), st))
            pending.append(self._exchange_p2p(f, False) if p2p else self._exchange(ks[f], xs[f], False))
        self._mark("y_fwd")
        for f in range(no):
            pending[f].wait()
            self._check(lib.ddl_slab_zfwd(h, 1, _ptrs([ks[f]]), _ptrs([b["e"][f]]), 0, st))
        self._mark("wait+z_fwd")
        self._assemble(physics_id, pp, b["e"][:no], state, deriv, fuse, st)
        self._mark("assemble")
