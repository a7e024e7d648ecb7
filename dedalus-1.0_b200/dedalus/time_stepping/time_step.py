"""Time integrators (reference: dedalus/time_stepping/time_step.py).

RK2mid / RK2trap follow the reference's ETD2RK schemes exactly; the per-component Cython loops
(forward_step_cy_{2d,3d}.pyx) are replaced by ONE multi-component device kernel per stage that
rebuilds the integrating factor from the wavenumbers in registers (include/ddl.h: ddl_stage).

RK4 and CrankNicholsonVisc cannot run in the reference as shipped (time_step.py:209,214 call
undefined linear_step / intfac_step; :449 passes the same StateData as input and output;
:495 uses a removed one-argument RHS API).  They are implemented here as RESTATED in
SURVEY.md section 8(c): classical RK4 data flow of :426-483 with distinct k buffers and
forward_step = Euler where the integrating factor is None, etd1(-IF) otherwise; CN as
y+ = ((1/dt - IF/2) y + N(y)) / (1/dt + IF/2).
"""
import ctypes as C
import os
import pickle
import time

import numpy as np

from .. import _lib
from .._lib import lib, check
from ..data_objects import plan as _plan
from ..utils.logger import mylog
from ..utils.parallelism import com_sys
from ..utils.timer import timer

hg_version = "b200-native"


def _kspace_tensors(sd):
    """The k-space buffers of every component (each brought to k-space first); the list itself is cached on the StateData."""
    comps, ks = sd._cached()[:2]
    for c in comps:
        if c._curr_space != "kspace":
            c.require_space("kspace")
    return ks


def _kspace_ptrs(sd):
    """(void*)[n] of those buffers as ctypes passes it: built once, the buffers never change identity."""
    comps, _, arr, ptr, _ = sd._cached()
    for c in comps:
        if c._curr_space != "kspace":
            c.require_space("kspace")
    return arr, ptr


def _all_clean(*sds):
    """True when every component of every StateData is known to vanish outside the dealias mask."""
    for sd in sds:
        if sd is not None:
            for c in sd._cached()[0]:
                if not c._clean:
                    return False
    return True


def _mark(sd, clean):
    """Record what a kernel of ours left in `sd`: zero outside the mask, or (computed from operands that were not) content
    there -- a known status either way, so no device check is needed before the next step."""
    for c in sd._cached()[0]:
        c._clean = clean
        if not clean:
            c._checked = True


def _inherit(out, start):
    """A stage update out = a * start + (combination of derivatives) keeps the solenoidal-or-not verdict of
    its start state: derivatives of u are projected, those of B are curls (physics.verify_solenoidal)."""
    if out is start:
        return
    for co, cs in zip(out._cached()[0], start._cached()[0]):
        co._soln = cs._soln


def _settle(sd, R=None):
    """Before a step: components whose buffer was handed out since the last step get their
    'dealiased' bit re-checked on the device (FourierRepresentation.verify_clean) and, given the physics
    object, their fields' 'solenoidal' verdict (Physics.verify_solenoidal): both decide which kernels run."""
    if R is not None and getattr(R, "_unfused", False):
        return                      # shearing box: no fused kernels to choose between
    for c in sd._cached()[0]:
        if c._escaped:
            c.refresh_escaped()     # the caller kept the tensor and wrote to it since the last step?
    if R is not None and hasattr(R, "sync_knowledge"):
        R.sync_knowledge(sd)        # slab runs: ranks agree on which buffers need a collective check (no-op on one rank)
        R._sync_done_for = id(sd)   # the RHS evaluation of this state that follows need not ask again
    todo = None
    for c in sd._cached()[0]:
        if c._curr_space != "kspace":
            c.require_space("kspace")
        if not c._clean and not c._checked:
            todo = True
    if todo:
        from ..data_objects.representations import verify_clean_many
        verify_clean_many(sd._cached()[0])          # one device pass and one host read for the whole state
    if R is not None and hasattr(R, "verify_solenoidal"):
        R.verify_solenoidal(sd)


def _plan_of(sd):
    return sd._cached()[0][0]._plan


def _if_coefficients(deriv):
    """(coeff[ncomp], order) from the integrating factors the physics attached to `deriv`."""
    co, order = [], 1
    for _, _, c in deriv.components():
        IF = c.integrating_factor
        if IF is None:
            co.append(0.0)
        else:
            co.append(IF.coeff)
            order = IF.order
    return (C.c_double * len(co))(*co), order


class TimeStepBase(object):
    """Stopping controls, statistics and the snapshot trigger (time_step.py:48-179)."""

    timer = timer

    def __init__(self, RHS, CFL=0.1, int_factor=None):
        self.RHS = RHS
        self.CFL = CFL
        self.int_factor = int_factor
        self.sim_stop_time = 1000
        self.wall_stop_time = 60. * 60.
        self.stop_iteration = 1000
        self.save_cadence = 100
        self.max_save_period = 100.
        self.time = 0
        self.iteration = 0
        self._nsnap = 0
        self._tlastsnap = 0.
        self.dt_old = np.finfo("d").max / 10.
        self._start_time = time.time()

    @property
    def ok(self):
        if self.iteration >= self.stop_iteration:
            why = "stop iteration reached."
        elif self.time >= self.sim_stop_time:
            why = "simulation stop time reached."
        elif (time.time() - self._start_time) >= self.wall_stop_time:
            why = "wall stop time reached."
        else:
            return True
        if com_sys.myproc == 0:
            mylog.info("Timestepping complete: " + why)
        return False

    @timer
    def advance(self, data, dt=None):
        if (self.iteration % self.save_cadence) == 0 or (self.time - self._tlastsnap >= self.max_save_period):
            self.snapshot(data)
        if dt is None and not (self.fuse_cfl and self._lazy_dt_ok()):
            dt = self.cfl_dt(data)
        self.do_advance(data, dt)          # dt None: the step takes its CFL limit from its own first RHS evaluation
        mylog.info("step %i" % self.iteration)

    def do_advance(self, data, dt):
        raise NotImplementedError("do_advance must be provided by subclass.")

    # ---- CUDA-graph replay of a fixed-dt step (not in the reference) ------------------------
    def do_advance_graph(self, data, dt):
        """do_advance(data, dt) replayed from a CUDA graph: the kernel sequence of one step is
        captured once per (state, dt) and re-launched with a single driver call.  For the small
        2-D grids (BASELINE configs 1-2: tens of launches of a few microseconds each) the step is
        launch-bound and this removes the launch overhead; large grids gain nothing.  Single-rank
        plans only; anything that changes the launch sequence (a different dt, handing a buffer
        to the caller, which drops its 'dealiased' bit) re-captures."""
        import torch
        comps = [c for _, _, c in data.components()]
        if comps[0]._plan.nranks != 1:
            return self.do_advance(data, dt)
        for c in comps:
            c.require_space("kspace")
            c.refresh_escaped()
        key = (id(data), float(dt), tuple(c._clean for c in comps))
        g = getattr(self, "_graph", None)
        if g is None or g[0] != key:
            # the first two steps of a (state, dt) run eagerly: lazy set-up (workspace, function
            # attributes, integrating factors) and the switch to the fused stage path must happen
            # outside the capture
            warm = getattr(self, "_graph_warm", None)
            if warm is None or warm[0] != (id(data), float(dt)):
                warm = self._graph_warm = [(id(data), float(dt)), 0]
            if warm[1] < 2:
                warm[1] += 1
                return self.do_advance(data, dt)
            t0, it0, dt0 = self.time, self.iteration, data.time
            graph = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                with torch.cuda.graph(graph, stream=s):
                    self.do_advance(data, dt)
            torch.cuda.current_stream().wait_stream(s)
            clean_after = tuple(c._clean for c in comps)
            # the capture does not execute: undo its host-side bookkeeping, then replay it once
            self.time, self.iteration = t0, it0
            data.set_time(dt0)
            if clean_after != key[2]:
                # the step itself changes the state flags: the next step would launch a different
                # sequence, so this one is not replayable; run it eagerly
                for c, cl in zip(comps, key[2]):
                    c._clean = cl
                self._graph = None
                return self.do_advance(data, dt)
            self._graph = g = (key, graph, clean_after)
        g[1].replay()
        for c, cl in zip(comps, g[2]):
            c._clean = cl
        data.set_time(data.time + dt)
        self.time += dt
        self.iteration += 1

    def __getstate__(self):
        """Pickled into every snapshot (time_step.py:133-138): the integrating-factor coefficients are a ctypes
        array at run time and travel as a plain list."""
        state = dict(self.__dict__)
        state.pop("_fuse_cache", None)          # argument blocks of the fused stage calls: rebuilt on demand
        state.pop("_graph", None)
        state.pop("_graph_warm", None)
        co = state.get("_coeff")
        if co is not None:
            state["_coeff"] = ("coeff", list(co[0]), int(co[1]))
        return state

    def __setstate__(self, state):
        co = state.get("_coeff")
        if isinstance(co, tuple) and co and co[0] == "coeff":
            state["_coeff"] = ((C.c_double * len(co[1]))(*co[1]), co[2])
        self.__dict__.update(state)

    @timer
    def snapshot(self, data):
        """Per-rank snapshot directory snap_%05i with the reference's contents (time_step.py:112-151): the source of
        the forcing functions (forcing_functions.py), the pickled physics / state-layout / integrator objects
        (dedalus_obj_%04i.cpkl; the field arrays are not in the pickle) and the fields.  Field file: HDF5 with the
        reference's layout -- /time, attribute hg_version, /fields/<name>/<comp> with a 'space' attribute
        (fields.py:118-125) -- when h5py is importable; otherwise one .npy per component plus fields.cpu%04i.json
        naming the space each one was saved in.  dedalus.utils.restart.restart() reads both."""
        import inspect
        import json
        rank = com_sys.myproc
        path = "snap_%05i" % self._nsnap
        if rank == 0 and not os.path.exists(path):
            os.mkdir(path)
        if com_sys.comm:
            com_sys.comm.barrier()
        source = ""
        for k, forcer in self.RHS.forcing_functions.items():
            if forcer:
                try:
                    source += inspect.getsource(forcer) + "\n"
                except (OSError, TypeError):
                    pass                    # defined interactively: the pickle keeps only its name
                self.RHS._forcing_function_names[k] = forcer.__name__
        if rank == 0:
            with open(os.path.join(path, "forcing_functions.py"), "w") as f:
                f.write(source)
        with open(os.path.join(path, "dedalus_obj_%04i.cpkl" % rank), "wb") as f:
            pickle.dump(self.RHS, f)
            pickle.dump(data, f)
            pickle.dump(self, f)
        try:
            import h5py
        except ImportError:
            h5py = None
        if h5py is not None:
            with h5py.File(os.path.join(path, "data.cpu%04i" % rank), mode="w") as out:
                out.create_dataset("time", data=self.time)
                out.attrs["hg_version"] = hg_version
                data.snapshot(out.create_group("/fields"))
        else:
            index = {"time": float(self.time), "hg_version": hg_version, "fields": {}}
            for name, i, c in data.components():
                fn = "%s_%i.cpu%04i.npy" % (name, i, rank)
                buf = c._k if c._curr_space == "kspace" else c.xdata      # internal buffers: saving does not hand them out
                np.save(os.path.join(path, fn), buf.cpu().numpy())
                index["fields"]["%s/%i" % (name, i)] = {"file": fn, "space": c._curr_space}
            with open(os.path.join(path, "fields.cpu%04i.json" % rank), "w") as f:
                json.dump(index, f)
        self._nsnap += 1
        self._tlastsnap = self.time

    def final_stats(self):
        self.timer.print_stats()
        if com_sys.myproc == 0:
            total = self.timer.timers.get("advance", 0.0)
            print("total advance wall time: %10.5e sec" % total)
            print("%10.5e sec/step " % (total / max(self.iteration, 1)))
            print()
            print("Simulation complete. Status: awesome")

    def finalize(self, data):
        self.snapshot(data)
        self.final_stats()

    def cfl_dt(self, data):
        return self._limit_dt(self.CFL * self.RHS.compute_dt(data))

    def _limit_dt(self, dt):
        if dt > 1.05 * self.dt_old:      # at most 5% growth per step
            dt = 1.05 * self.dt_old
        self.dt_old = dt
        mylog.info("dt = %10.5e" % dt)
        return dt

    # ---- CFL limit taken from the step's own first RHS evaluation ---------------------------
    # advance(data) in the reference is compute_dt(data) -- every component of u and B to x-space and
    # back, 12 + 12 transforms in 3-D MHD -- followed by do_advance.  The first RHS evaluation of a
    # step is AT that very state and has u(x), B(x) in registers inside its x pass, so the maxima are
    # reduced there (include/ddl.h: ddl_rhs_capture_max): do_advance(data, None) evaluates k1 unfused
    # (no dt needed), reads the two maxima back, fixes dt and carries on with the fused stages.
    fuse_cfl = True         # set False for the reference's route (tests compare both)

    def _lazy_dt_ok(self):
        R = self.RHS
        return hasattr(R, "capture_begin") and not R.aux_eqns and not getattr(R, "_unfused", False)

    def _rhs_and_dt(self, data, k):
        """k = RHS(data) and the CFL-limited dt of `data`, from the same x pass."""
        R = self.RHS
        token = R.capture_begin(data)
        try:
            R.RHS(data, k)
        finally:
            maxima = R.capture_end(token)
        return self._limit_dt(self.CFL * R.dt_from_maxima(data, maxima))

    # ---- stage update fused with the spectral assembly of the RHS --------------------------
    fuse_stages = True      # set False to force the unfused RHS + stage-kernel path (tests compare both)

    def _can_fuse(self, derivs, states):
        """ddl_rhs_stage applies: nothing else touches deriv (rotation, forcing, aux equations), the
        integrating-factor coefficients are known (after the first unfused step), the state is solenoidal (else the
        advective-form RHS, which has no fused stage) and every operand vanishes outside the dealias mask (the fused sweep
        visits the retained modes only).  The one exception to the last rule: `states` (the evolving state and the stage
        states built from it) of a physics without linear terms out there may carry content outside the mask (hydro never
        dealiases its state) -- _stage_fused then updates those entries with ddl_stage_outside."""
        R = self.RHS
        if not self.fuse_stages or getattr(self, "_coeff", None) is None or not hasattr(R, "_fused_rhs"):
            return False
        if getattr(R, "_unfused", False):
            return False
        if R.aux_eqns or not R.can_fuse_stage():
            return False
        for sd in derivs:
            for c in sd._cached()[0]:
                if not c._soln or not c._clean:
                    return False
        junk_ok = getattr(R, "_junk_keeps_fused", False)
        for sd in states:
            for c in sd._cached()[0]:
                if not c._soln:
                    return False
                if not c._clean and not (junk_ok and c._checked):
                    return False
        return True

    def _stage_fused(self, kind, state_in, start, out, dt_step, deriv1=None, k_out=None, total=None, wdiv=1., first=0, last=0):
        """out = stage(kind)(start, RHS(state_in)[, deriv1]) with the derivative consumed in registers
        (include/ddl.h: ddl_rhs_stage); k_out: also store it (a later stage needs it).  The argument block of a call site
        (pointer lists, the ddl_stage_fuse struct) is built once and reused: only the step sizes change from call to call."""
        R = self.RHS
        key = (kind, id(state_in), id(start), id(out), id(deriv1), id(k_out), id(total), first, last)
        cache = self.__dict__.setdefault("_fuse_cache", {})
        ent = cache.get(key)
        if ent is None or ent[1] is not self._coeff:
            coeff, order = self._coeff
            opt = [None if sd is None else _kspace_ptrs(sd)[1] for sd in (total, deriv1, k_out)]
            fuse = _lib.StageFuse(_kspace_ptrs(start)[1], opt[0], _kspace_ptrs(out)[1], C.cast(coeff, C.c_void_p),
                                  int(order), int(first), int(last), float(wdiv), float(dt_step), int(kind), 0, opt[1], opt[2])
            marks = [sd for sd in (out, total, k_out) if sd is not None]
            # the StateData objects are kept alive with the entry: their ids are the key
            ent = cache[key] = (fuse, self._coeff, marks, (state_in, start, out, deriv1, k_out, total))
        fuse, _, marks, _ = ent
        fuse.wdiv, fuse.dt_step = float(wdiv), float(dt_step)
        for sd in (start, out, total, deriv1, k_out):
            if sd is not None:
                _kspace_tensors(sd)                 # everything in k-space (no-ops in steady state)
        R._fused_rhs(state_in, None, R._rhs_flags(), fuse=fuse)
        junk = not _all_clean(start)
        if junk and not (kind == _lib.ETD2RK1 and out is start):
            # content outside the mask (_can_fuse admitted it): the integrating factor alone acts there
            coeff, order = self._coeff
            check(lib.ddl_stage_outside(_plan_of(start).handle, int(kind), len(start._cached()[0]), _kspace_ptrs(start)[0],
                                        _kspace_ptrs(out)[0], coeff, int(order), float(dt_step), _plan.current_stream()))
        for sd in marks:
            _mark(sd, True)
        if junk:
            for c in out._cached()[0]:
                c._clean, c._checked = False, True
        _inherit(out, start)
        if k_out is not None:
            k_out.set_time(state_in.time)

    # ---- device stage launches ---------------------------------------------------------
    def _stage_tensor(self, kind, start, out, d1, d2, if_from, dt):
        """The same updates with the integrating factor as an ARRAY (shearing box: k^2 moves with time), component by
        component like the reference's loops (time_step.py:285-304,372-386): one launch of the array-factor kernel per
        component (include/ddl.h: ddl_step_array), Euler forms where the factor is None."""
        s, o, a = _kspace_tensors(start), _kspace_tensors(out), _kspace_tensors(d1)
        b = _kspace_tensors(d2) if d2 is not None else [None] * len(s)
        ndim = s[0].dim()
        for j, (_, _, c) in enumerate(if_from.components()):
            IF = c.integrating_factor
            neg = None if IF is None else (-IF.tensor()).contiguous()
            check(lib.ddl_step_array(kind, ndim, s[j].numel(), s[j].data_ptr(), o[j].data_ptr(), a[j].data_ptr(),
                                     b[j].data_ptr() if b[j] is not None else None, neg.data_ptr() if neg is not None else None,
                                     float(dt), _plan.current_stream()))
        _mark(out, False)

    def _stage(self, kind, start, out, d1, d2, if_from, dt):
        if getattr(self.RHS, "_dynamic_k", False):
            return self._stage_tensor(kind, start, out, d1, d2, if_from, dt)
        s, o, a = _kspace_ptrs(start)[0], _kspace_ptrs(out)[0], _kspace_ptrs(d1)[0]
        b = _kspace_ptrs(d2)[0] if d2 is not None else None
        coeff, order = _if_coefficients(if_from)
        clean = _all_clean(start, out, d1, d2)
        check(lib.ddl_stage(_plan_of(start).handle, kind, len(s), s, o, a, b, coeff, order,
                            float(dt), _lib.STAGE_RETAINED_ONLY if clean else 0, _plan.current_stream()))
        _mark(out, clean)
        _inherit(out, start)


class RKBase(TimeStepBase):
    """Base class for the Runge-Kutta integrators."""

    @timer
    def forward_step(self, start, deriv, output, dt):
        """output = start advanced by dt with `deriv`: forward Euler for components without an
        integrating factor, first-order ETD otherwise (time_step.py:187-221; the undefined
        linear_step / intfac_step are euler / etd1(-IF))."""
        self._stage(_lib.ETD1, start, output, deriv, None, deriv, dt)
        output.set_time(start.time + dt)


class RK2mid(RKBase):
    """Second-order explicit midpoint RK with exponential time differencing (time_step.py:224-309)."""

    def __init__(self, *arg, **kwargs):
        TimeStepBase.__init__(self, *arg, **kwargs)
        self.data2 = self.RHS.create_fields(0.)
        self.data2._internal = True                     # written by our kernels only (Physics.sync_knowledge)
        self.deriv1 = self.RHS.create_fields(0.)
        self.deriv2 = self.RHS.create_fields(0.)

    def do_advance(self, data, dt):
        _settle(data, self.RHS)
        data2, k1, k2 = self.data2, self.deriv1, self.deriv2
        lazy = dt is None
        if not lazy and self._can_fuse((k1,), (data, data2)):
            self._stage_fused(_lib.ETD1, data, data, data2, dt / 2., k_out=k1)              # k1 kept for the second stage
            data2.set_time(data.time + dt / 2.)
            self._stage_fused(_lib.ETD2RK2, data2, data, data, dt, deriv1=k1)               # k2 never stored
            data.set_time(data.time + dt)
            self.time += dt
            self.iteration += 1
            return
        if lazy:
            dt = self._rhs_and_dt(data, k1)
        else:
            self.RHS.RHS(data, k1)
        if getattr(self, "_coeff", None) is None:
            self._coeff = _if_coefficients(k1)
        self._stage(_lib.ETD1, data, data2, k1, None, k1, dt / 2.)        # a_n (euler where IF is None)
        data2.set_time(data.time + dt / 2.)
        if lazy and self._can_fuse((k1,), (data, data2)):
            self._stage_fused(_lib.ETD2RK2, data2, data, data, dt, deriv1=k1)
        else:
            self.RHS.RHS(data2, k2)
            self._stage(_lib.ETD2RK2, data, data, k1, k2, k1, dt)
        data.set_time(data.time + dt)
        self.time += dt
        self.iteration += 1


class RK2trap(RKBase):
    """Second-order explicit trapezoidal RK with ETD (time_step.py:312-392)."""

    def __init__(self, *arg, **kwargs):
        TimeStepBase.__init__(self, *arg, **kwargs)
        self.deriv1 = self.RHS.create_fields(0.)
        self.deriv2 = self.RHS.create_fields(0.)

    def do_advance(self, data, dt):
        _settle(data, self.RHS)
        k1, k2 = self.deriv1, self.deriv2
        lazy = dt is None
        if not lazy and self._can_fuse((k1,), (data,)):
            self._stage_fused(_lib.ETD1, data, data, data, dt, k_out=k1)
            data.set_time(data.time + dt)
            self._stage_fused(_lib.ETD2RK1, data, data, data, dt, deriv1=k1)
            self.time += dt
            self.iteration += 1
            return
        if lazy:
            dt = self._rhs_and_dt(data, k1)
        else:
            self.RHS.RHS(data, k1)
        if getattr(self, "_coeff", None) is None:
            self._coeff = _if_coefficients(k1)
        self._stage(_lib.ETD1, data, data, k1, None, k1, dt)
        data.set_time(data.time + dt)
        if lazy and self._can_fuse((k1,), (data,)):
            self._stage_fused(_lib.ETD2RK1, data, data, data, dt, deriv1=k1)
        else:
            self.RHS.RHS(data, k2)
            self._stage(_lib.ETD2RK1, data, data, k1, k2, k1, dt)
        self.time += dt
        self.iteration += 1


class RK4(RKBase):
    """Classical fourth-order RK (tableau of time_step.py:399-411), restated -- see module doc."""

    def __init__(self, *arg, **kwargs):
        TimeStepBase.__init__(self, *arg, **kwargs)
        self.temp_data = self.RHS.create_fields(0.)     # stage states
        self.temp_data._internal = True                 # written by our kernels only (Physics.sync_knowledge)
        self.total_deriv = self.RHS.create_fields(0.)   # (k1 + 2 k2 + 2 k3 + k4) / 6
        self.k_data = self.RHS.create_fields(0.)        # current k_i
        self._coeff = None

    def _rk4(self, y, out, wdiv, dt_step, first, last):
        pl = _plan_of(y)
        ys, ks = _kspace_ptrs(y)[0], _kspace_ptrs(self.k_data)[0]
        ts, os_ = _kspace_ptrs(self.total_deriv)[0], _kspace_ptrs(out)[0]
        coeff, order = self._coeff
        clean = _all_clean(y, self.k_data, self.total_deriv, out)
        # what the sweep leaves outside the mask: total = [total +] k / w is zero there when its operands are; out = S(y, .) when y is too
        tot_clean = _all_clean(self.k_data) and (bool(first) or _all_clean(self.total_deriv))
        out_clean = tot_clean and _all_clean(y)
        check(lib.ddl_rk4_stage(pl.handle, len(ys), ys, ks, ts, os_, coeff, order, float(wdiv), float(dt_step), int(first), int(last),
                                _lib.STAGE_RETAINED_ONLY if clean else 0, _plan.current_stream()))
        _mark(self.total_deriv, tot_clean)
        _mark(out, out_clean)
        _inherit(out, y)

    def _rk4_fused(self, state_in, y, out, wdiv, dt_step, first, last):
        self._stage_fused(_lib.FUSE_RK4, state_in, y, out, dt_step, total=self.total_deriv, wdiv=wdiv, first=first, last=last)

    def _advance_dynamic_k(self, data, dt):
        """Shearing box (restated like the rest of RK4; no run of the reference can pin it): every stage step takes the
        integrating factor of the container holding its derivative -- each RHS evaluation rewrites it at its own time level
        (physics.py:584-586) -- and the final combination the copy taken at k1 (time_step.py:437-441).  Array-factor kernel per
        component (include/ddl.h: ddl_step_array), accumulation as tensor operations."""
        if dt is None:
            dt = self.cfl_dt(data)
        R, tmp, k, tot = self.RHS, self.temp_data, self.k_data, self.total_deriv
        ndim = data.ndim

        def neg_factors(sd):
            return [None if c.integrating_factor is None else (-c.integrating_factor.tensor()).contiguous() for _, _, c in sd.components()]

        def step(deriv, out, h, factors):
            s, d, o = _kspace_tensors(data), _kspace_tensors(deriv), _kspace_tensors(out)
            for j in range(len(s)):
                f = factors[j]
                check(lib.ddl_step_array(_lib.ETD1, ndim, s[j].numel(), s[j].data_ptr(), o[j].data_ptr(), d[j].data_ptr(), None,
                                         f.data_ptr() if f is not None else None, float(h), _plan.current_stream()))
            _mark(out, False)

        R.RHS(data, k)                                            # k1
        initial = neg_factors(k)
        for t, kk in zip(_kspace_tensors(tot), _kspace_tensors(k)):
            t.copy_(kk / 6.)
        step(k, tmp, dt / 2., initial)
        tmp.set_time(data.time + dt / 2.)
        for w, h in ((3., dt / 2.), (3., dt), (6., None)):
            R.RHS(tmp, k)                                         # k2, k3, k4
            for t, kk in zip(_kspace_tensors(tot), _kspace_tensors(k)):
                t.add_(kk / w)
            if h is not None:
                step(k, tmp, h, neg_factors(k))
                tmp.set_time(data.time + h)
        step(tot, data, dt, initial)
        data.set_time(data.time + dt)
        self.time += dt
        self.iteration += 1

    def _advance_fused(self, data, dt):
        tmp = self.temp_data
        self._rk4_fused(data, data, tmp, 6., dt / 2., True, False)      # k1
        tmp.set_time(data.time + dt / 2.)
        self._rk4_fused(tmp, data, tmp, 3., dt / 2., False, False)      # k2
        self._rk4_fused(tmp, data, tmp, 3., dt, False, False)           # k3
        tmp.set_time(data.time + dt)
        self._rk4_fused(tmp, data, data, 6., dt, False, True)           # k4 ; y+ = S(y, total + k4/6, dt)
        data.set_time(data.time + dt)
        self.time += dt
        self.iteration += 1

    def do_advance(self, data, dt):
        R, tmp, k = self.RHS, self.temp_data, self.k_data
        if getattr(R, "_dynamic_k", False):
            return self._advance_dynamic_k(data, dt)
        _settle(data, self.RHS)
        lazy = dt is None
        if not lazy and self._can_fuse((self.total_deriv,), (data, self.temp_data)):
            return self._advance_fused(data, dt)
        aux = list(R.aux_eqns.values())
        a_old = [a.value for a in aux]
        a_final = [a.RHS(a.value) / 6. for a in aux]
        if lazy:
            dt = self._rhs_and_dt(data, k)                 # k1 and the CFL limit of y from the same x pass
        else:
            R.RHS(data, k)                                 # k1
        if self._coeff is None:
            self._coeff = _if_coefficients(k)
            for (_, _, ct), (_, _, ck) in zip(self.total_deriv.components(), k.components()):
                ct.integrating_factor = ck.integrating_factor
        self._rk4(data, tmp, 6., dt / 2., True, False)     # total = k1/6 ; tmp = S(y, k1, dt/2)
        tmp.set_time(data.time + dt / 2.)
        if lazy and self._can_fuse((self.total_deriv,), (data, tmp)):
            self._rk4_fused(tmp, data, tmp, 3., dt / 2., False, False)      # k2
            self._rk4_fused(tmp, data, tmp, 3., dt, False, False)           # k3
            tmp.set_time(data.time + dt)
            self._rk4_fused(tmp, data, data, 6., dt, False, True)           # k4
            data.set_time(data.time + dt)
            self.time += dt
            self.iteration += 1
            return
        for j, a in enumerate(aux):
            a.value = a_old[j] + dt / 2. * a.RHS(a.value)
            a_final[j] += a.RHS(a.value) / 3.
        R.RHS(tmp, k)                                      # k2
        self._rk4(data, tmp, 3., dt / 2., False, False)
        for j, a in enumerate(aux):
            a.value = a_old[j] + dt / 2. * a.RHS(a.value)
            a_final[j] += a.RHS(a.value) / 3.
        R.RHS(tmp, k)                                      # k3
        self._rk4(data, tmp, 3., dt, False, False)
        tmp.set_time(data.time + dt)
        for j, a in enumerate(aux):
            a.value = a_old[j] + dt * a.RHS(a.value)
            a_final[j] += a.RHS(a.value) / 6.
        R.RHS(tmp, k)                                      # k4
        self._rk4(data, data, 6., dt, False, True)         # y+ = S(y, total + k4/6, dt)
        data.set_time(data.time + dt)
        for j, a in enumerate(aux):
            a.value = a_old[j] + dt * a_final[j]
        self.time += dt
        self.iteration += 1


class CrankNicholsonVisc(TimeStepBase):
    """Crank-Nicholson on the viscous term, explicit nonlinear term (time_step.py:486-506), restated."""

    def __init__(self, *arg, **kwargs):
        TimeStepBase.__init__(self, *arg, **kwargs)
        self.deriv = self.RHS.create_fields(0.)
        self._coeff = None

    def do_advance(self, data, dt):
        if getattr(self.RHS, "_dynamic_k", False):
            # shearing box (restated): the factor of this RHS evaluation's time level, as tensor operations
            if dt is None:
                dt = self.cfl_dt(data)
            self.RHS.RHS(data, self.deriv)
            for (_, _, c), y, kk in zip(self.deriv.components(), _kspace_tensors(data), _kspace_tensors(self.deriv)):
                IF = 0. if c.integrating_factor is None else c.integrating_factor.tensor()
                top, bottom = 1. / dt - 0.5 * IF, 1. / dt + 0.5 * IF
                y.copy_(top / bottom * y + kk / bottom)
            _mark(data, False)
            data.set_time(data.time + dt)
            self.time += dt
            self.iteration += 1
            return
        _settle(data, self.RHS)
        lazy = dt is None
        if not lazy and self._can_fuse((), (data,)):
            self._stage_fused(_lib.FUSE_CN, data, data, data, dt)
            data.set_time(data.time + dt)
            self.time += dt
            self.iteration += 1
            return
        if lazy:
            dt = self._rhs_and_dt(data, self.deriv)
        else:
            self.RHS.RHS(data, self.deriv)
        if self._coeff is None:
            self._coeff = _if_coefficients(self.deriv)
        ys, ks = _kspace_tensors(data), _kspace_tensors(self.deriv)
        coeff, order = self._coeff
        clean = _all_clean(data, self.deriv)
        check(lib.ddl_cn_step(_plan_of(data).handle, len(ys), _lib.ptr_array(ys), _lib.ptr_array(ks), coeff, order,
                              float(dt), _lib.STAGE_RETAINED_ONLY if clean else 0, _plan.current_stream()))
        _mark(data, clean)
        data.set_time(data.time + dt)
        self.time += dt
        self.iteration += 1
