"""ctypes driver for the HOST-EMULATION build of the CUDA library (tests only).

The emulation build (dedalus-1.0_b200/build.py --emul) compiles the very same kernel bodies
with g++; one host thread walks every work item.  It exists so that the index logic of the
kernels can be checked against the oracle in the GPU-less build container.  The product
package never loads it.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BUILD = os.path.join(ROOT, "tests", "host", "_build")


class PhysParams(C.Structure):
    _fields_ = [("rho0", C.c_double), ("g", C.c_double), ("alpha_t", C.c_double), ("beta", C.c_double),
                ("boussinesq_dir", C.c_int), ("reserved", C.c_int)]


def load():
    sys.path.insert(0, os.path.join(ROOT, "dedalus-1.0_b200"))
    import build as ddl_build
    if os.environ.get("DDL_TEST_HOST_EMUL_ASAN") == "1":      # tests/test_host_sanitizer.py: instrumented build, libasan preloaded
        lib = C.CDLL(ddl_build.build_emul(BUILD + "_asan", sanitize=True))
    else:
        lib = C.CDLL(ddl_build.build_emul(BUILD))
    lib.ddl_last_error.restype = C.c_char_p
    lib.ddl_version.restype = C.c_char_p
    lib.ddl_workspace_bytes.restype = C.c_size_t
    lib.ddl_rhs_workspace_bytes.restype = C.c_size_t
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _pa(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


class EmulPlan:
    PHYS = {"IncompressibleHydro": 0, "BoussinesqHydro": 1, "IncompressibleMHD": 2}
    ADV = 3          # include/ddl.h DDL_*_ADV = base + 3: advective-form policies for non-solenoidal states

    def __init__(self, lib, grid, nranks=1, rank=0, layout=0):
        """grid: oracle Grid (supplies k values and the dealias mask exactly as the host layer will).
        nranks > 1: plan of rank `rank` of a slab decomposition (k arrays stay global)."""
        self.lib, self.g = lib, grid
        nd = grid.ndim
        shape = np.array(grid.shape, dtype=np.int64)
        k = {n: np.ascontiguousarray(v.ravel(), dtype=np.float64) for n, v in grid.k.items()}
        mask = grid.dealias_mask()
        keep = {}
        for name in k:
            axis = grid.ktrans[name]
            other = tuple(a for a in range(nd) if a != axis)
            keep[name] = np.ascontiguousarray((~mask).any(axis=other).astype(np.uint8))
        self._keep = keep
        self._k = k
        self.plan = C.c_void_p()
        rc = lib.ddl_plan_create_slab(C.byref(self.plan), nd, _p(shape), _p(k["x"]), _p(k["y"]),
                                      _p(k["z"]) if nd == 3 else None, _p(keep["x"]), _p(keep["y"]),
                                      _p(keep["z"]) if nd == 3 else None, nranks, rank, layout)
        self.check(rc)
        self.work = None

    def check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.ddl_last_error().decode())

    def _ws(self, nbytes):
        if self.work is None or self.work.nbytes < nbytes:
            self.work = np.zeros(nbytes // 16 + 1, dtype=np.complex128)
        return self.work

    def forward(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        k = np.full(self.g.kshape, np.nan + 0j, dtype=np.complex128)
        w = self._ws(self.lib.ddl_workspace_bytes(self.plan, 1, 1))
        self.check(self.lib.ddl_forward(self.plan, _p(x), _p(k), _p(w), C.c_size_t(w.nbytes), None))
        return k

    def backward(self, k):
        k = np.ascontiguousarray(k, dtype=np.complex128).copy()
        x = np.full(self.g.shape, np.nan, dtype=np.float64)
        w = self._ws(self.lib.ddl_workspace_bytes(self.plan, 1, 1))
        self.check(self.lib.ddl_backward(self.plan, _p(k), _p(x), _p(w), C.c_size_t(w.nbytes), None))
        return x, k

    def deriv(self, k, axis):
        k = np.ascontiguousarray(k, dtype=np.complex128)
        o = np.empty_like(k)
        self.check(self.lib.ddl_deriv(self.plan, _p(k), _p(o), axis, None))
        return o

    def rhs(self, physics, params, state, flags=1, adv=False):
        """adv: use the advective-form policy (state need not be solenoidal); the divergence scratch arrays the
        ABI wants after the state are appended here."""
        pid = self.PHYS[physics] + (self.ADV if adv else 0)
        pp = PhysParams(params.get("rho0", 1.0), params.get("g", 1.0), params.get("alpha_t", 1.0),
                        params.get("beta", 1.0), {"x": 0, "y": 1, "z": 2}[params.get("boussinesq_direction", "z")], 0)
        state = [np.ascontiguousarray(s, dtype=np.complex128).copy() for s in state]
        deriv = [np.full(state[0].shape, np.nan + 0j, dtype=np.complex128) for _ in state]
        scratch = [np.full(state[0].shape, np.nan + 0j, dtype=np.complex128) for _ in range(2 if physics == "IncompressibleMHD" else 1)] if adv else []
        w = self._ws(self.lib.ddl_rhs_workspace_bytes(self.plan, pid))
        self.check(self.lib.ddl_rhs(self.plan, pid, C.byref(pp), _pa(state + scratch), _pa(deriv), _p(w), C.c_size_t(w.nbytes),
                                    flags, None))
        return np.stack(deriv), np.stack(state)

    def rhs_stage(self, physics, params, state, kind, y, coeff, vo, dt, total=None, deriv1=None, want_k=False, wdiv=1.0,
                  first=0, last=0, flags=0):
        """ddl_rhs_stage: returns dict(out=..., total=..., k=...) after the fused assembly + stage update."""
        class Fuse(C.Structure):
            _fields_ = [("y", C.c_void_p), ("total", C.c_void_p), ("out", C.c_void_p), ("coeff", C.c_void_p),
                        ("visc_order", C.c_int), ("first", C.c_int), ("last", C.c_int), ("wdiv", C.c_double), ("dt_step", C.c_double),
                        ("kind", C.c_int), ("reserved", C.c_int), ("deriv1", C.c_void_p), ("k_out", C.c_void_p)]
        pid = self.PHYS[physics]
        pp = PhysParams(params.get("rho0", 1.0), params.get("g", 1.0), params.get("alpha_t", 1.0),
                        params.get("beta", 1.0), {"x": 0, "y": 1, "z": 2}[params.get("boussinesq_direction", "z")], 0)
        cp = lambda arrs: [np.ascontiguousarray(s, dtype=np.complex128).copy() for s in arrs]
        state, y = cp(state), cp(y)
        out = [np.zeros_like(s) for s in state]
        co = np.ascontiguousarray(coeff, dtype=np.float64)
        w = self._ws(self.lib.ddl_rhs_workspace_bytes(self.plan, pid))
        keep = [_pa(y), _pa(out)]
        res = {}
        opt = []
        for name, arrs in (("total", total), ("deriv1", deriv1), ("k", [np.zeros_like(s) for s in state] if want_k else None)):
            if arrs is None:
                opt.append(None)
            else:
                arrs = cp(arrs)
                res[name] = arrs
                pa = _pa(arrs)
                keep.append(pa)
                opt.append(C.cast(pa, C.c_void_p))
        fu = Fuse(C.cast(keep[0], C.c_void_p), opt[0], C.cast(keep[1], C.c_void_p), _p(co), vo, int(first), int(last),
                  float(wdiv), float(dt), int(kind), 0, opt[1], opt[2])
        self.check(self.lib.ddl_rhs_stage(self.plan, pid, C.byref(pp), _pa(state), _p(w), C.c_size_t(w.nbytes), flags,
                                          C.byref(fu), None))
        res["out"] = out
        return {k: np.stack(v) for k, v in res.items()}

    def rk4_stage(self, y, k, total, coeff, vo, wdiv, dt, first, last, flags=1):
        y = [np.ascontiguousarray(s, dtype=np.complex128).copy() for s in y]
        total = [np.ascontiguousarray(s, dtype=np.complex128).copy() for s in total]
        out = [np.zeros_like(s) for s in y]
        co = np.ascontiguousarray(coeff, dtype=np.float64)
        self.check(self.lib.ddl_rk4_stage(self.plan, len(y), _pa(y), _pa(list(k)), _pa(total), _pa(out), _p(co), vo,
                                          C.c_double(wdiv), C.c_double(dt), int(first), int(last), flags, None))
        return np.stack(out), np.stack(total)

    def cn_step(self, y, k, coeff, vo, dt, flags=1):
        y = [np.ascontiguousarray(s, dtype=np.complex128).copy() for s in y]
        co = np.ascontiguousarray(coeff, dtype=np.float64)
        self.check(self.lib.ddl_cn_step(self.plan, len(y), _pa(y), _pa(list(k)), _p(co), vo, C.c_double(dt), flags, None))
        return np.stack(y)

    # ---- reductions (include/ddl.h: ddl_reduce_invariants / ddl_reduce_max_square / ddl_rhs_capture_max)
    NINV = 24

    def _pp(self, params):
        return PhysParams(params.get("rho0", 1.0), params.get("g", 1.0), params.get("alpha_t", 1.0),
                          params.get("beta", 1.0), {"x": 0, "y": 1, "z": 2}[params.get("boussinesq_direction", "z")], 0)

    def invariants(self, physics, state, flags=0):
        state = [np.ascontiguousarray(s, dtype=np.complex128) for s in state]
        out = np.full(self.NINV, np.nan)
        self.check(self.lib.ddl_reduce_invariants(self.plan, self.PHYS[physics], _pa(state), flags, _p(out), None))
        return out

    def max_square(self, physics, params, state, flags=0):
        pid = self.PHYS[physics]
        pp = self._pp(params)
        state = [np.ascontiguousarray(s, dtype=np.complex128).copy() for s in state]
        out = np.full(2, np.nan)
        w = self._ws(self.lib.ddl_rhs_workspace_bytes(self.plan, pid))
        self.check(self.lib.ddl_reduce_max_square(self.plan, pid, C.byref(pp), _pa(state), _p(w), C.c_size_t(w.nbytes),
                                                  flags, _p(out), None))
        return out, np.stack(state)

    def capture_max(self, out2):
        """out2: float64[2] the next RHS evaluations maximise into, or None to switch the capture off."""
        self._cap = out2
        self.check(self.lib.ddl_rhs_capture_max(self.plan, _p(out2) if out2 is not None else None))

    def stage(self, kind, start, d1, d2, coeff, vo, dt, flags=0):
        n = len(start)
        out = [np.zeros_like(s) for s in start]
        co = np.ascontiguousarray(coeff, dtype=np.float64)
        self.check(self.lib.ddl_stage(self.plan, kind, n, _pa(start), _pa(out), _pa(d1), _pa(d2) if d2 is not None else None,
                                      _p(co), vo, C.c_double(dt), flags, None))
        return out
