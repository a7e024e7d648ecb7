"""Physics classes: IncompressibleHydro, BoussinesqHydro, IncompressibleMHD
(reference: dedalus/physics/physics.py).

Same constructor, parameters dictionary and ``RHS(data, deriv)`` contract as the reference, but
the right-hand side is ONE fused device pipeline (include/ddl.h: ddl_rhs) instead of ~100 numpy
passes and 15-36 separate FFTs: inverse transforms of the state, the quadratic products at
every grid point inside the middle transform pass (the real-space fields never reach HBM),
forward transforms of the products, then derivatives / curl / solenoidal projection in one
spectral assembly kernel.  The conservative forms used are algebraically identical to the
reference's advective forms for a dealiased, solenoidal state.

The unfused vector-calculus helpers of the reference (XgradY, XcrossY, curlX, divX, ...) are
kept as thin tensor-level methods for analysis scripts; the hot path does not use them.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .._lib import lib, check
from ..config import decfg
from ..data_objects.api import create_field_classes, AuxEquation, StateData
from ..data_objects import plan as _plan
from ..utils.logger import mylog
from ..utils.parallelism import com_sys
from ..utils.timer import timer


class IntegratingFactor(object):
    """c * (k^2)**order, kept symbolic: the stage kernels rebuild it from integer wavenumbers in
    registers, so no N_k-sized array exists unless somebody asks for one
    (reference: one real array per component, physics.py:504-522)."""

    def __init__(self, comp, coeff, order):
        self.comp, self.coeff, self.order = comp, float(coeff), int(order)

    def tensor(self):
        return self.coeff * self.comp.k2() ** self.order

    def __neg__(self):
        return -self.tensor()

    def __mul__(self, other):
        return self.tensor() * other

    __rmul__ = __mul__

    def copy(self):
        return IntegratingFactor(self.comp, self.coeff, self.order)


def _reconstruct_object(cls, state):
    obj = cls(state["shape"], state["_representation"], state["length"])
    obj.__dict__.update(state)
    return obj


class Physics(object):
    """Base class: defines fields and provides a right-hand side for the integrators."""

    _physics_id = None

    def __init__(self, shape, representation, length=None):
        self.shape = shape
        self._representation = representation
        self.length = length if length else (2 * np.pi,) * len(shape)
        self.ndim = len(self.shape)
        self.dims = range(self.ndim)
        self._field_list = []
        self._aux_field_list = []
        self.parameters = {}
        self.aux_eqns = {}
        self.forcing_functions = {}
        self._forcing_function_names = {}
        self._field_classes = create_field_classes(self._representation, self.shape, self.length)
        self._is_finalized = False
        self._tracer = decfg.getboolean("physics", "use_tracer")
        self.k2 = None
        self._trans = {0: "x", 1: "y", 2: "z"}

    def __getitem__(self, item):
        value = self.parameters.get(item, None)
        if value is None:
            raise KeyError
        return value

    def __reduce__(self):
        self._is_finalized = False
        keep = {k: v for k, v in self.__dict__.items() if k not in ("aux_fields", "_field_classes", "forcing_functions", "_pp_cache",
                                                                    "_aux_fields")}
        return (_reconstruct_object, (self.__class__, keep))

    def _finalize(self):
        self._is_finalized = True

    def create_fields(self, time, field_list=None):
        if field_list is None:
            field_list = self._field_list
        return StateData(time, self.shape, self.length, self._field_classes, field_list=field_list,
                         params=self.parameters)

    @property
    def aux_fields(self):
        """Scratch fields of the reference (mathscalar, mathvector, ...).  The fused RHS needs
        none of them; they are built on first access for scripts that use the helpers."""
        if getattr(self, "_aux_fields", None) is None:
            self._aux_fields = self.create_fields(0., self._aux_field_list)
        return self._aux_fields

    def _setup_aux_eqns(self, aux_eqns, RHS, ics, kwarglists):
        for f, r, ic, kwargs in zip(aux_eqns, RHS, ics, kwarglists):
            self.aux_eqns[f] = AuxEquation(r, kwargs, ic)

    def compute_dt(self, data):
        """Raw time-step limit without the CFL number (physics.py:151-158).

        The reference takes every component of u (and B) to x-space and back for this (6 + 6
        transforms per call in 3-D MHD, fields.py:153-157); here the maxima come out of ONE pass of
        the inverse half of the RHS pipeline with the reduction inside the x pass
        (include/ddl.h: ddl_reduce_max_square), and nothing is written but workspace."""
        if self._unfused:
            self.dtlist = []
            self.set_dtlist(data)          # field.max_square(), component by component (fields.py:153-157)
            return min(self.dtlist)
        return self.dt_from_maxima(data, self.max_squares(data))

    def dt_from_maxima(self, data, maxima):
        """min(dtlist) with max_square of u / B already known: `maxima` = (max u_i^2, max B_i^2) as
        returned by max_squares() or captured inside an RHS evaluation (capture_begin/_end)."""
        self.dtlist = []
        self._maxsq = maxima
        try:
            self.set_dtlist(data)
        finally:
            self._maxsq = None
        return min(self.dtlist)

    # ------------------------------------------------------------------ CFL maxima on the device
    _maxsq = None

    def _field_max_square(self, data, fname, slot):
        """data[fname].max_square(): from the fused reduction inside compute_dt, else the
        reference's component-by-component route (fields.py:153-157)."""
        if self._maxsq is not None:
            return self._maxsq[slot]
        return data[fname].max_square()

    def max_squares(self, data):
        """(max_{x,i} u_i^2, max_{x,i} B_i^2 [T^2; 0 for hydro]) of `data`, reduced over ranks."""
        import torch
        from ..utils.parallelism import reduce_max
        if not self._is_finalized:
            self._finalize()
        data = self._main(data)
        self.sync_knowledge(data)
        state = self._max_square_side_effects(data)
        pl = next(data.components())[2]._plan
        out = torch.zeros(2, dtype=torch.float64, device=pl.device)
        pp = self._phys_params()
        if pl.nranks > 1:
            pl.pipeline.max_square(self._physics_id, pp, state, out, False)
        else:
            w = pl.rhs_workspace(self._physics_id)
            check(lib.ddl_reduce_max_square(pl.handle, self._physics_id, C.byref(pp), _lib.ptr_array(state), w.data_ptr(),
                                            w.numel(), 0, out.data_ptr(), _plan.current_stream()))
        return self._finish_maxima(out)

    @staticmethod
    def _max_square_side_effects(data):
        """What the reference's max_square leaves behind: u and B (not T) have been through x-space and
        back (fields.py:153-157), so their spectra are the dealiased spectra of real fields, in place
        (representations.py:347-357).  No-ops for buffers our own kernels produced.  Returns the k tensors."""
        state = []
        for name, _, c in data.components():
            c.require_space("kspace")
            if name in ("u", "B"):
                c._hermitian_project()
                if not c._clean:
                    c.dealias()
            state.append(c._k)
        return state

    @staticmethod
    def _finish_maxima(out):
        """(max u^2, max B^2) over all ranks from the two device values: one all-reduce, one 16-byte host read."""
        from ..utils.parallelism import com_sys
        if com_sys.comm is not None:
            import torch.distributed as dist
            out = out.clone()
            dist.all_reduce(out, op=dist.ReduceOp.MAX)
        a, b = out.cpu().tolist()
        return (a, b)

    def capture_begin(self, data):
        """Ask the x passes of the following RHS evaluations to maximise u_i(x)^2 / B_i(x)^2 into a
        device buffer (include/ddl.h: ddl_rhs_capture_max): the time-step limit of the state an RHS
        is evaluated at then costs no transform at all.  Returns the token for capture_end()."""
        import torch
        data = self._main(data)
        self.sync_knowledge(data)
        self._max_square_side_effects(data)
        pl = next(data.components())[2]._plan
        out = torch.zeros(2, dtype=torch.float64, device=pl.device)
        check(lib.ddl_rhs_capture_max(pl.handle, out.data_ptr()))
        return (pl, out)

    def capture_end(self, token):
        """Switch the capture off and return the maxima (one 16-byte device -> host read)."""
        pl, out = token
        check(lib.ddl_rhs_capture_max(pl.handle, None))
        return self._finish_maxima(out)

    # ------------------------------------------------------------------ fused RHS
    def _phys_params(self):
        p = self.parameters
        key = (p.get("rho0"), p.get("g"), p.get("alpha_t"), p.get("beta"), p.get("boussinesq_direction"), self._tracer)
        hit = self.__dict__.get("_pp_cache")
        if hit is not None and hit[0] == key:
            return hit[1]
        pp = self._build_phys_params()
        self._pp_cache = (key, pp)
        return pp

    def _build_phys_params(self):
        p = self.parameters
        d = p.get("boussinesq_direction", "z")
        if self._tracer and not self._tracer_extra:
            return _lib.PhysParams(float(p.get("rho0", 1.0)), 0.0, 0.0, 0.0, 0, 0)      # tracer in the T slot, no coupling
        return _lib.PhysParams(float(p.get("rho0", 1.0)), float(p.get("g", 1.0)), float(p.get("alpha_t", 1.0)),
                               float(p.get("beta", 1.0)), {"x": 0, "y": 1, "z": 2}[d], 0)

    def RHS(self, data, deriv):
        """Base behaviour of the reference (physics.py:134-149): zero deriv, synchronise times."""
        if not self._is_finalized:
            self._finalize()
        for _, field in deriv:
            field.zero_all()
        deriv.set_time(data.time)
        if not self._representation._static_k:
            # shearing box: the scratch fields' wavenumbers follow the state's time too (physics.py:143-149)
            for _, field in self.aux_fields:
                field.zero_all(space="kspace")
            self.aux_fields.set_time(data.time)

    @property
    def _dynamic_k(self):
        """Wavenumbers that move with time (shearing box): the stage kernels' static tables do not apply either."""
        return not self._representation._static_k

    @property
    def _unfused(self):
        """True for representations whose wavenumbers depend on time (FourierShearRepresentation): the fused
        pipeline's index tables are static, so the right-hand side is evaluated the reference's way, helper by
        helper (a compatibility path), and the integrators update with tensor operations."""
        flag = self.__dict__.get("_unfused_flag")
        if flag is None:
            # decided once per physics object (the representation reads FFT.dealiasing when it is constructed, too): this
            # property sits on the per-step path of the launch-bound 2-D configurations
            if not self._representation._static_k:
                flag = True
            else:
                # without the per-axis 2/3 rule the reference's x-space products are ALIASED and its advective forms are what
                # they are: the fused pipeline (conservative products, transforms pruned to the retained modes) has no
                # equivalent, the helper sequence reproduces it exactly
                flag = decfg.get("FFT", "dealiasing") not in ("2/3", "2/3 cython")
            self._unfused_flag = flag
        return flag

    # ------------------------------------------------------------------ solenoidal or not
    SOLENOIDAL_TOL = 1e-12      # compressive fraction sqrt(sum |k.u|^2 / sum |k|^2 |u|^2) below which u counts as div-free

    # ------------------------------------------------------------------ passive tracer on BoussinesqHydro / IncompressibleMHD
    @property
    def _tracer_extra(self):
        """True when this class carries the tracer next to fields of its own (T or B): the state then is (u, c, T | B)."""
        return self._tracer and type(self).__name__ != "IncompressibleHydro"

    def _substate(self, sd, names, replace=None):
        """A StateData over the SAME field objects as `sd`, restricted to `names` (cached on sd): what the fused kernels of
        this class (state without 'c') and the tracer pass (u, c) take.  `replace` swaps in other field objects by name."""
        key = tuple(names) + (id(replace.get("u")) if replace else 0,)
        subs = sd.__dict__.setdefault("_subs", {})      # not pickled (StateData.__reduce__)
        sub = subs.get(key)
        if sub is None:
            sub = sd.clone()
            for n in names:
                sub.fields[n] = (replace or {}).get(n) or sd.fields[n]
            sub._cache = None
            sub.__dict__["_internal"] = sd.__dict__.get("_internal", False)
            subs[key] = sub
        sub.time = sd.time
        return sub

    def _main(self, sd):
        """`sd` without the tracer field (everything this class's own kernels and reductions see)."""
        if not self._tracer_extra or "c" not in sd.fields:
            return sd
        return self._substate(sd, [n for n in sd.fields if n != "c"])

    def _tracer_pass(self, data, deriv):
        """deriv['c'] = -u.grad c through a plain hydro-with-tracer object (the Boussinesq kernels with g = alpha_t = beta = 0
        and c in the T slot, as IncompressibleHydro itself does); the velocity derivative it also produces goes to scratch."""
        eng = self.__dict__.get("_tracer_engine")
        if eng is None:
            old = decfg.get("physics", "use_tracer")
            decfg.set("physics", "use_tracer", "True")
            try:
                eng = IncompressibleHydro(self.shape, self._representation, self.length)
            finally:
                decfg.set("physics", "use_tracer", old)
            eng._finalize()
            self._tracer_engine = eng
            self._tracer_scratch = deriv.clone()
            self._tracer_scratch.add_field("u", "VectorField")
        trc_d = self._substate(data, ["u", "c"])
        trc_k = self._substate(deriv, ["u", "c"], replace={"u": self._tracer_scratch.fields["u"]})
        eng._fused_rhs(trc_d, trc_k, _lib.RHS_ZERO_FILL)

    def sync_knowledge(self, data):
        """Slab-decomposed runs: make what the ranks know about `data`'s buffers the same on every rank BEFORE anything
        branches on it.  `_soln` / `_sym` are dropped on the rank that hands a buffer out only (kdata, comp['kspace']),
        and callers do touch buffers asymmetrically (init_cond.taylor_green / alfven write on the ranks find_mode finds;
        rank 0 prints a mode) -- while the checks those bits gate are device collectives: the invariants sweep
        (all_reduce) and the Hermitian projection of the kx = 0 plane (all_gather).  One host-side OR per call (com_sys.host_or)
        of [some vector component unverified, component i not known Hermitian]; every rank then adopts the union, so
        all of them enter the same collectives.  Integrator-internal stage states (written by our kernels only) skip it."""
        comps = data.comp_list()
        if comps[0]._plan.nranks == 1 or data.__dict__.get("_internal"):
            return
        if self.__dict__.pop("_sync_done_for", None) == id(data):
            return                         # _settle() of this very step has just done it for this state
        bits = 0
        for i, c in enumerate(comps):
            if c._soln is None:
                bits |= 1
            if not c._sym:
                bits |= 2 << i
        union = com_sys.host_or(bits)
        if union == 0:
            return
        vector = set(id(c) for n, f in data if n in ("u", "B") for _, c in f)
        for i, c in enumerate(comps):
            if (union & 1) and id(c) in vector:
                c._soln = None             # some rank must re-measure: everyone takes part in the sweep
            if union & (2 << i):
                c._sym = False             # rows of the plane changed somewhere: every rank re-projects its rows

    def verify_solenoidal(self, data):
        """Make sure every vector component of `data` knows whether its field is solenoidal (`_soln`).

        The fused pipeline's conservative products equal the reference's advective forms only for
        div u = div B = 0 (include/ddl.h DDL_*_ADV).  Buffers our own kernels wrote inherit the answer
        (derivatives are projected / curls, so an update keeps whatever its start state had); for a buffer
        the CALLER has written since the last check one sweep of ddl_reduce_invariants measures the
        compressive fraction of u and B.  Returns True when the whole state is solenoidal."""
        if self._unfused:
            return True                    # the unfused path evaluates the advective forms themselves
        data = self._main(data)
        for c in data.comp_list():
            if c._soln is not True:
                break
        else:
            return True                    # steady state of a run: every verdict known and positive
        vec = [(n, f) for n, f in data if n in ("u", "B")]
        for n, f in data:
            if n not in ("u", "B"):
                for _, c in f:
                    c._soln = True         # scalars (T, tracer) have no divergence to speak of: never hold the fused path back
        if any(c._soln is None for _, f in vec for _, c in f):
            from ..analysis.volume_average import invariants
            from .._lib import INV
            # over the retained modes only: content outside the mask (hydro never dealiases its state; the
            # Nyquist-row entries of the reference's 2-D Taylor-Green field) never enters a product of the pipeline
            inv = invariants(data, retained_only=True)
            if inv is None:
                raise NotImplementedError("verify_solenoidal: not a standard u [+ T | B] state")
            tot = inv[0]
            for n, f in vec:
                comp, rot = (tot[INV["div2"]], 2 * tot[INV["enstrophy"]]) if n == "u" else (tot[INV["mag_div2"]], 2 * tot[INV["current2"]])
                ok = bool(comp <= self.SOLENOIDAL_TOL ** 2 * (comp + rot))
                for _, c in f:
                    c._soln = ok
        return all(c._soln for _, f in vec for _, c in f)

    def can_fuse_stage(self):
        """True when nothing is added to deriv after the fused pipeline (rotation, forcing), so an
        integrator may ask for the spectral assembly fused with its stage update."""
        if not self._is_finalized:
            self._finalize()
        if self._unfused or self._tracer_extra:
            return False
        return not getattr(self, "_rotation", False) and not self.forcing_functions

    def _fused_rhs(self, data, deriv, flags, fuse=None):
        """deriv = RHS(data).  With `fuse` (an _lib.StageFuse; deriv is None) the derivative is
        consumed in registers by the integrator's stage update instead of being written (ddl_rhs_stage)."""
        if not self._is_finalized:
            self._finalize()
        if self._unfused:
            raise NotImplementedError(
                "The fused RHS uses conservative products, which equal the reference's advective form only under "
                "2/3 dealiasing; FFT.dealiasing=%r is not supported." % decfg.get("FFT", "dealiasing"))
        if self._tracer_extra and "c" in data.fields:
            if fuse is not None:
                raise RuntimeError("the fused stage update is not available with a tracer next to T / B (can_fuse_stage)")
            self._tracer_pass(data, deriv)
            data, deriv = self._main(data), self._main(deriv)
        comps = data.comp_list()
        for c in comps:
            if c._escaped:
                c.refresh_escaped()        # written through a tensor the caller kept? (representations.py)
        self.sync_knowledge(data)
        state_clean = deriv_clean = True
        for c in comps:
            if c._curr_space != "kspace":
                c.require_space("kspace")
            if (flags & _lib.RHS_DEALIAS_STATE) and not c._sym:
                c._hermitian_project()      # with the mask: the image of the reference's x-space round trip of the state
            state_clean = state_clean and c._clean
        state = data._cached()[1]
        out = []
        if deriv is not None:
            for c in deriv.comp_list():
                c._curr_space = "kspace"
                deriv_clean = deriv_clean and c._clean
            out = deriv._cached()[1]
        # mask passes only where the buffers are not already known to be zero outside the mask
        if deriv_clean:
            flags &= ~_lib.RHS_ZERO_FILL
        if state_clean:
            flags &= ~_lib.RHS_DEALIAS_STATE
        pl = comps[0]._plan
        pp = self._phys_params()
        if not self.verify_solenoidal(data):
            # a compressive part in u or B (whatever put it there, the reference keeps it): advective-form
            # policies, which need scratch arrays for the divergence spectra after the state
            if fuse is not None:
                raise RuntimeError("the fused stage update needs a solenoidal state (time_step._can_fuse checks it)")
            return self._advective_rhs(pl, pp, data, deriv, state, out, flags)
        if pl.nranks > 1:
            # slab-decomposed: pipeline phases with the exchange between the z and y passes
            pl.pipeline.rhs(self._physics_id, pp, state, out, bool(flags & _lib.RHS_DEALIAS_STATE),
                            bool(flags & _lib.RHS_ZERO_FILL) and fuse is None, fuse=fuse)
        elif fuse is not None:
            w = pl.rhs_workspace(self._physics_id)
            check(lib.ddl_rhs_stage(pl.handle, self._physics_id, C.byref(pp), data._cached()[3], w.data_ptr(), w.numel(),
                                  flags & ~_lib.RHS_ZERO_FILL, C.byref(fuse), _plan.current_stream()))
        else:
            w = pl.rhs_workspace(self._physics_id)
            check(lib.ddl_rhs(pl.handle, self._physics_id, C.byref(pp), data._cached()[3], deriv._cached()[3],
                              w.data_ptr(), w.numel(), flags, _plan.current_stream()))
        if deriv is not None:
            for c in deriv.comp_list():
                c._clean = True
                c._soln = True
        if flags & _lib.RHS_DEALIAS_STATE:
            for c in comps:
                c._clean = True
        if deriv is not None:
            deriv.set_time(data.time)

    def _advective_rhs(self, pl, pp, data, deriv, state, out, flags):
        import torch
        pid = self._physics_id + _lib.ADV
        nth = 2 if self._physics_id == _lib.MHD else 1
        scratch = getattr(pl, "_theta_scratch", None)
        if scratch is None or len(scratch) < nth:
            scratch = pl._theta_scratch = [torch.empty_like(state[0]) for _ in range(nth)]
        ptrs = state + scratch[:nth]
        if pl.nranks > 1:
            pl.pipeline.rhs(pid, pp, ptrs, out, bool(flags & _lib.RHS_DEALIAS_STATE), bool(flags & _lib.RHS_ZERO_FILL),
                            ncomp=len(state))
        else:
            w = pl.rhs_workspace(pid)
            check(lib.ddl_rhs(pl.handle, pid, C.byref(pp), _lib.ptr_array(ptrs), _lib.ptr_array(out), w.data_ptr(), w.numel(),
                              flags, _plan.current_stream()))
        for _, _, c in deriv.components():
            c._clean = True
            c._soln = True
        if flags & _lib.RHS_DEALIAS_STATE:
            for _, _, c in data.components():
                c._clean = True
        deriv.set_time(data.time)

    # ------------------------------------------------------------------ unfused helpers
    def gradX(self, X, output):
        """output[N*i + j] = d X_i / d x_j (physics.py:160-178)."""
        N = self.ndim
        for i in range(X.ncomp):
            for j in self.dims:
                output[N * i + j]["kspace"] = X[i].deriv(self._trans[j])

    def divX(self, X, output):
        acc = 0
        for i in range(X.ncomp):
            acc = acc + X[i].deriv(self._trans[i])
        output["kspace"] = acc

    def XgradY(self, X, Y, stmp, vtmp, output):
        """(X . grad) Y evaluated in x-space, component by component (physics.py:197-228)."""
        xs = [X[i]["kspace"].clone() for i in range(X.ncomp)]
        for i in range(X.ncomp):
            vtmp[i]["kspace"] = xs[i]
        for ci in range(Y.ncomp):
            acc = 0
            for i in self.dims:
                stmp["kspace"] = Y[ci].deriv(self._trans[i])
                acc = acc + stmp["xspace"] * vtmp[i]["xspace"]
            output[ci]["xspace"] = acc

    def XconstcrossY(self, X, Y, output):
        Yk = [Y[i]["kspace"] for i in range(Y.ncomp)]
        if self.ndim == 2:
            if np.isscalar(X):
                output["x"]["kspace"] = -X * Yk[1]
                output["y"]["kspace"] = X * Yk[0]
            else:
                Xy, Xx = X
                output["kspace"] = Xx * Yk[1] - Xy * Yk[0]
        else:
            Xz, Xy, Xx = X
            output["x"]["kspace"] = Xy * Yk[2] - Xz * Yk[1]
            output["y"]["kspace"] = Xz * Yk[0] - Xx * Yk[2]
            output["z"]["kspace"] = Xx * Yk[1] - Xy * Yk[0]

    def XcrossY(self, X, Y, output):
        """X x Y in x-space (physics.py:309-354); like the reference this leaves its inputs in x-space."""
        Yx = [Y[i]["xspace"] for i in range(Y.ncomp)]
        if self.ndim == 2:
            if X.ncomp == 1:
                Xz = X["xspace"]
                output["x"]["xspace"] = -Xz * Yx[1]
                output["y"]["xspace"] = Xz * Yx[0]
            else:
                output["xspace"] = X["x"]["xspace"] * Yx[1] - X["y"]["xspace"] * Yx[0]
        else:
            Xx = [X[i]["xspace"] for i in range(3)]
            output["x"]["xspace"] = Xx[1] * Yx[2] - Xx[2] * Yx[1]
            output["y"]["xspace"] = Xx[2] * Yx[0] - Xx[0] * Yx[2]
            output["z"]["xspace"] = Xx[0] * Yx[1] - Xx[1] * Yx[0]

    def XdotY(self, X, Y, output, space):
        if X.ncomp != Y.ncomp:
            raise ValueError("Vectors not the same size")
        acc = 0
        for i in range(X.ncomp):
            acc = acc + X[i][space] * Y[i][space]
        output[space] = acc

    def curlX(self, X, output):
        t = self._trans
        if X.ncomp == 3:
            output[0]["kspace"] = X[2].deriv(t[1]) - X[1].deriv(t[2])
            output[1]["kspace"] = X[0].deriv(t[2]) - X[2].deriv(t[0])
            output[2]["kspace"] = X[1].deriv(t[0]) - X[0].deriv(t[1])
        elif X.ncomp == 2:
            output["kspace"] = X[1].deriv(t[0]) - X[0].deriv(t[1])
        else:
            output[0]["kspace"] = X[0].deriv(t[1])
            output[1]["kspace"] = -X[0].deriv(t[0])

    # physics.py:412-413 caches k^2 at the FIRST laplace_solve and keeps it -- also in a shearing box, where the wavenumbers
    # drift, so that the reference's pressure solve divides by the k^2 of that first time level (the swinging-wave sample then
    # departs from Lithwick's analytic solution, tests/test_gpu_known_answers.py).  True reproduces the reference (the
    # default: drop-in parity); False re-evaluates k^2 at every solve.
    cache_k2 = True

    @timer
    def laplace_solve(self, X, output):
        """Solve laplace(output) = X (physics.py:407-416)."""
        if self.k2 is None or not self.cache_k2:
            self.k2 = output.k2(no_zero=True)
        output["kspace"] = -X["kspace"] / self.k2


class IncompressibleHydro(Physics):
    """Homogeneous incompressible hydrodynamics (physics.py:419-610).

    parameters: 'nu' (0.), 'viscosity_order' (1), 'shear_rate' (0.; nonzero needs
    FourierShearRepresentation), 'Omega' (None)."""

    _physics_id = _lib.HYDRO

    def __init__(self, *args, **kwargs):
        Physics.__init__(self, *args, **kwargs)
        self._field_list.append(("u", "VectorField"))
        self._aux_field_list.append(("mathscalar", "ScalarField"))
        self._aux_field_list.append(("mathvector", "VectorField"))
        self.parameters["viscosity_order"] = 1
        self.parameters["nu"] = 0.
        self.parameters["shear_rate"] = 0.
        self.parameters["Omega"] = None
        if self._tracer:
            # passive scalar c advected by u (physics.py:468-470,516-522,580-582): d_t c = -u.grad c + c_diff lap c.
            # That is the Boussinesq temperature equation without buoyancy and stratification, so the plain
            # hydro class runs the Boussinesq kernels with g = alpha_t = beta = 0 and c in the T slot.
            # BoussinesqHydro / IncompressibleMHD inherit the tracer (physics.py:467-470: field 'c' right after 'u'): their own
            # kernels run on the state without 'c', and an internal plain-hydro-with-tracer object evaluates d_t c from
            # (u, c) in a second pass (_tracer_pass; its velocity derivative goes to scratch).
            self._field_list.append(("c", "ScalarField"))
            self.parameters["c_diff"] = 0.
            if type(self) is IncompressibleHydro:
                self._physics_id = _lib.BOUSSINESQ
        self._first_rhs = True

    def __reduce__(self):
        self._first_rhs = True
        return Physics.__reduce__(self)

    def _finalize(self):
        Physics._finalize(self)
        # physics.py:486-494
        if self.parameters["shear_rate"] == 0.:
            self._shear = False
            if self._unfused:
                mylog.warning("Performance suffers when using a shearing representation without a linear shear.")
        else:
            self._shear = True
            if not self._unfused:
                raise ValueError("A shearing representation must be used if shear_rate is nonzero.")
        self._rotation = self.parameters["Omega"] is not None
        if self._rotation and self.ndim == 2:
            mylog.warning("Rotation is dynamically insignificant in 2D incompressible hydrodynamics.")

    def _setup_integrating_factors(self, deriv):
        nu, vo = self.parameters["nu"], self.parameters["viscosity_order"]
        for _, comp in deriv["u"]:
            comp.integrating_factor = None if nu == 0. else IntegratingFactor(comp, nu, vo)
        if self._tracer:
            comp, diff = deriv["c"][0], self.parameters["c_diff"]
            comp.integrating_factor = None if diff == 0. else IntegratingFactor(comp, diff, vo)

    def set_velocity_forcing(self, func):
        self.forcing_functions["VelocityForcing"] = func

    @property
    def _junk_keeps_fused(self):
        """A state with content outside the dealias mask may still take the fused stage kernel when nothing but the integrating
        factor acts out there: plain hydro (with or without the passive tracer).  Boussinesq has buoyancy / stratification on
        the full arrays (physics.py:691-708); MHD dealiases its state anyway."""
        return type(self) is IncompressibleHydro

    def _rhs_flags(self):
        return _lib.RHS_ZERO_FILL

    def RHS(self, data, deriv):
        """deriv = RHS(data):  d_t u = P[-u.grad u (+ class terms)]  (physics.py:527-599)."""
        if self._first_rhs:
            self._setup_integrating_factors(deriv)
            self._first_rhs = False
        if not self._is_finalized:
            self._finalize()
        if self._unfused:
            return self._unfused_hydro_rhs(data, deriv)
        self._fused_rhs(data, deriv, self._rhs_flags())
        self._extra_momentum_terms(data, deriv)

    def _unfused_hydro_rhs(self, data, deriv):
        """The reference's own sequence for the velocity (and tracer) equation, physics.py:527-599, in the form
        d_t f + S y d_x f - c lap f = RHS(f); used for the shearing box."""
        Physics.RHS(self, data, deriv)
        ms, mv = self.aux_fields["mathscalar"], self.aux_fields["mathvector"]
        S, Om = self.parameters["shear_rate"], self.parameters["Omega"]
        self.XgradY(data["u"], data["u"], ms, mv, deriv["u"])
        for i in self.dims:
            deriv["u"][i]["kspace"].mul_(-1.)
        if self._shear:
            deriv["u"]["x"]["kspace"].sub_(S * data["u"]["y"]["kspace"])
        if self._rotation:
            self.XconstcrossY(Om, data["u"], mv)
            for i in self.dims:
                deriv["u"][i]["kspace"].sub_(2 * mv[i]["kspace"])
        if "VelocityForcing" in self.forcing_functions:
            for i in self.dims:
                deriv["u"][i]["kspace"].add_(self.forcing_functions["VelocityForcing"](data, i))
        if type(self) is IncompressibleHydro:
            self._unfused_pressure_projection(data, deriv)
        if self._tracer:
            self.XgradY(data["u"], data["c"], ms, mv, deriv["c"])
            deriv["c"]["kspace"].mul_(-1.)
        if self._shear:
            self._setup_integrating_factors(deriv)      # k^2 has moved with the shear (physics.py:584-586)

    def _unfused_pressure_projection(self, data, deriv):
        """physics.py:588-599, including the shear source -S d_x u_y of the pressure equation and the k^2 array
        that laplace_solve caches at its FIRST call (physics.py:412-413): in a shearing box the reference keeps
        dividing by the k^2 of that first time level, and so does this."""
        ms = self.aux_fields["mathscalar"]
        self.divX(deriv["u"], ms)
        if self._shear:
            ms["kspace"].sub_(self.parameters["shear_rate"] * data["u"]["y"].deriv("x"))
        self.laplace_solve(ms, ms)
        for i in self.dims:
            deriv["u"][i]["kspace"].sub_(ms.deriv(self._trans[i]))

    def _extra_momentum_terms(self, data, deriv):
        """Rotation and user forcing (physics.py:563-572), added through the projector, which is
        linear: P[N + E] = P[N] + P[E].  Off the hot path (both default to off)."""
        extra = None
        if self._rotation:
            Om = self.parameters["Omega"]
            u = [data["u"][i]["kspace"] for i in self.dims]
            if self.ndim == 2:
                cross = [-Om * u[1], Om * u[0]]
            else:
                Oz, Oy, Ox = Om
                cross = [Oy * u[2] - Oz * u[1], Oz * u[0] - Ox * u[2], Ox * u[1] - Oy * u[0]]
            extra = [-2 * c for c in cross]
        if "VelocityForcing" in self.forcing_functions:
            f = [self.forcing_functions["VelocityForcing"](data, i) for i in self.dims]
            extra = f if extra is None else [a + b for a, b in zip(extra, f)]
        if extra is None:
            return
        c0 = deriv["u"][0]
        k = [c0.k[self._trans[i]] for i in self.dims]
        k2 = c0.k2(no_zero=True)
        kdot = sum(k[i] * extra[i] for i in self.dims) / k2
        for i in self.dims:
            deriv["u"][i]["kspace"].add_(extra[i] - k[i] * kdot)

    def pressure_projection(self, data, deriv):
        """Solenoidal projection of deriv['u'] as tensor operations (physics.py:588-599); the
        fused RHS already contains it."""
        c0 = deriv["u"][0]
        k = [c0.k[self._trans[i]] for i in self.dims]
        k2 = c0.k2(no_zero=True)
        kdot = sum(k[i] * deriv["u"][i]["kspace"] for i in self.dims) / k2
        for i in self.dims:
            deriv["u"][i]["kspace"].sub_(k[i] * kdot)

    def max_abs_vel(self, data):
        return np.sqrt(self._field_max_square(data, "u", 0))

    def set_dtlist(self, data):
        dx = data["u"]["x"].dx().min()
        self.dtlist.append(dx / self.max_abs_vel(data))


class BoussinesqHydro(IncompressibleHydro):
    """Boussinesq hydrodynamics (physics.py:612-721): adds T, 'kappa', 'g', 'alpha_t', 'beta'."""

    _physics_id = _lib.BOUSSINESQ

    def __init__(self, *args, **kwargs):
        IncompressibleHydro.__init__(self, *args, **kwargs)
        self._field_list.append(("T", "ScalarField"))
        self.parameters["kappa"] = 0.
        self.parameters["g"] = 1.
        self.parameters["alpha_t"] = 1.
        self.parameters["beta"] = 1.
        self.parameters["boussinesq_direction"] = decfg.get("physics", "boussinesq_direction")

    def _setup_integrating_factors(self, deriv):
        IncompressibleHydro._setup_integrating_factors(self, deriv)
        kappa, vo = self.parameters["kappa"], self.parameters["viscosity_order"]
        comp = deriv["T"][0]
        comp.integrating_factor = None if kappa == 0. else IntegratingFactor(comp, kappa, vo)

    def _linear_terms_outside_mask(self, data, deriv):
        """The reference adds buoyancy (g alpha_t T on u_dir, then the projection) and stratification
        (-beta u_dir on T) to the FULL k-space arrays (physics.py:691-708), so a state with content
        outside the 2/3 mask (hydro-type physics never dealias their state; e.g. the Nyquist-row
        entries sin_k / cos_k write, init_cond.py:137-143) feels them there too, while the nonlinear terms
        vanish there.  The fused pipeline covers the retained modes; this adds the masked-out ones.  Off the
        hot path: only for a state the caller has written modes outside the mask into (tensor operations)."""
        import torch
        c0 = deriv["u"][0]
        pl = c0._plan
        outside = getattr(pl, "_outside_mask", None)
        if outside is None:
            keep = None
            for name, kp in pl.keep_np.items():
                i = pl.ktrans[name]
                if i == 0 and pl.nranks > 1:
                    kp = kp[pl.krows]
                shp = [1] * pl.ndim
                shp[i] = len(kp)
                t = torch.from_numpy(np.ascontiguousarray(kp)).to(pl.device).reshape(shp)
                keep = t if keep is None else keep & t
            outside = pl._outside_mask = ~keep
        p = self.parameters
        d = {"x": 0, "y": 1, "z": 2}[p["boussinesq_direction"]]
        zero = torch.zeros((), dtype=torch.complex128, device=pl.device)
        buoy = torch.where(outside, (p["g"] * p["alpha_t"]) * data["T"].components[0]._k, zero)
        k = [c0.k[self._trans[i]] for i in self.dims]
        kdot = k[d] * buoy / c0.k2(no_zero=True)
        for i in self.dims:
            ci = deriv["u"][i]
            ci._k.sub_(k[i] * kdot)
            if i == d:
                ci._k.add_(buoy)
            ci._clean = False
            ci._checked = False
        dT = deriv["T"].components[0]
        dT._k.sub_(torch.where(outside, p["beta"] * data["u"][d]._k, zero))
        dT._clean = False
        dT._checked = False

    def set_thermal_forcing(self, func):
        self.forcing_functions["ThermalForcing"] = func

    def _unfused_bouss_rhs(self, data, deriv):
        """physics.py:664-712 helper by helper (shearing box)."""
        IncompressibleHydro.RHS(self, data, deriv)
        ms, mv = self.aux_fields["mathscalar"], self.aux_fields["mathvector"]
        p = self.parameters
        d = p["boussinesq_direction"]
        ms["kspace"] = data["T"]["kspace"]
        ms["kspace"].mul_(p["g"])
        ms["kspace"].mul_(p["alpha_t"])
        deriv["u"][d]["kspace"].add_(ms["kspace"])
        if type(self) is BoussinesqHydro:
            self._unfused_pressure_projection(data, deriv)
        self.XgradY(data["u"], data["T"], ms, mv, deriv["T"])
        deriv["T"]["kspace"].mul_(-1.)
        ms["kspace"] = data["u"][d]["kspace"]
        ms["kspace"].mul_(p["beta"])
        deriv["T"]["kspace"].sub_(ms["kspace"])
        if "ThermalForcing" in self.forcing_functions:
            deriv["T"]["kspace"].add_(self.forcing_functions["ThermalForcing"](data))

    def RHS(self, data, deriv):
        if self._unfused:
            return self._unfused_bouss_rhs(data, deriv)
        junk = not all(c._clean for _, _, c in data.components())
        IncompressibleHydro.RHS(self, data, deriv)
        if junk:
            self._linear_terms_outside_mask(data, deriv)
        if "ThermalForcing" in self.forcing_functions:
            deriv["T"]["kspace"].add_(self.forcing_functions["ThermalForcing"](data))

    def set_dtlist(self, data):
        IncompressibleHydro.set_dtlist(self, data)
        p = self.parameters
        dx = data["u"]["x"].dx().min()
        self.dtlist.append(dx / np.sqrt(np.abs(p["alpha_t"] * p["g"] * p["beta"])))


class IncompressibleMHD(IncompressibleHydro):
    """Homogeneous incompressible MHD (physics.py:724-836): adds B, 'rho0' (1.), 'eta' (0.)."""

    _physics_id = _lib.MHD

    def __init__(self, *args, **kwargs):
        IncompressibleHydro.__init__(self, *args, **kwargs)
        self._field_list.append(("B", "VectorField"))
        self._aux_field_list.append(("mathvector2", "VectorField"))
        self.parameters["rho0"] = 1.
        self.parameters["eta"] = 0.

    def _setup_integrating_factors(self, deriv):
        IncompressibleHydro._setup_integrating_factors(self, deriv)
        eta, vo = self.parameters["eta"], self.parameters["viscosity_order"]
        for _, comp in deriv["B"]:
            comp.integrating_factor = None if eta == 0. else IntegratingFactor(comp, eta, vo)

    def RHS(self, data, deriv):
        if self._unfused:
            return self._unfused_mhd_rhs(data, deriv)
        IncompressibleHydro.RHS(self, data, deriv)

    def _unfused_mhd_rhs(self, data, deriv):
        """physics.py:770-819 helper by helper (shearing box): Lorentz force, projection, induction, S B_y e_x."""
        IncompressibleHydro.RHS(self, data, deriv)
        aux = self.aux_fields
        ms, mv, mv2 = aux["mathscalar"], aux["mathvector"], aux["mathvector2"]
        fpr = 4 * np.pi * self.parameters["rho0"]
        cur = ms if self.ndim == 2 else mv
        self.curlX(data["B"], cur)
        self.XcrossY(cur, data["B"], mv2)
        for i in self.dims:
            deriv["u"][i]["kspace"].add_(mv2[i]["kspace"] / fpr)
        if type(self) is IncompressibleMHD:
            self._unfused_pressure_projection(data, deriv)
        self.XcrossY(data["u"], data["B"], cur)
        self.curlX(cur, deriv["B"])
        if self._shear:
            deriv["B"]["x"]["kspace"].add_(self.parameters["shear_rate"] * data["B"]["y"]["kspace"])

    def _rhs_flags(self):
        # the reference transforms the MHD state itself to x-space and back (physics.py:797-815),
        # which zeroes its masked-out modes in place
        return _lib.RHS_ZERO_FILL | _lib.RHS_DEALIAS_STATE

    def max_alfven_speed(self, data):
        fpr = 4 * np.pi * self.parameters["rho0"]
        return np.sqrt(self._field_max_square(data, "B", 1) / fpr)

    def set_dtlist(self, data):
        IncompressibleHydro.set_dtlist(self, data)
        dx = data["u"]["x"].dx().min()
        self.dtlist.append(dx / self.max_alfven_speed(data))
