#!/usr/bin/env python
"""Host-side cost of one time step: the package's Python between two launches, measured with every library call replaced by a
no-op (runs in the GPU-less build container through the host-emulation harness of tests/conftest.py; test infrastructure).
The 2-D BASELINE configurations are launch-bound (tens of microseconds of kernels per step), so this is their floor in eager mode.
    python profiles/host_overhead.py"""
import os
import sys
import time

os.environ["DDL_TEST_HOST_EMUL"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: F401,E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
from dedalus.mods import IncompressibleHydro, IncompressibleMHD, BoussinesqHydro, FourierRepresentation, RK2mid, RK4  # noqa: E402
import dedalus._lib as L  # noqa: E402
import dedalus.time_stepping.time_step as ts  # noqa: E402
import dedalus.physics.physics as ph  # noqa: E402


class Stub(object):
    def __init__(self, lib, calls):
        self._lib, self._calls = lib, calls

    def __getattr__(self, name):
        f = getattr(self._lib, name)
        if name == "ddl_last_error":
            return f

        def g(*a):
            self._calls[name] = self._calls.get(name, 0) + 1
            return 0
        return g


def run(physics, shape, integ, lazy_dt=False, n=2000):
    P = physics(shape, FourierRepresentation)
    P.parameters["nu"] = 1e-3
    data = P.create_fields(0.)
    rng = np.random.default_rng(0)
    for _, f in data:
        for _, c in f:
            c["xspace"] = torch.from_numpy(rng.standard_normal(shape))
            c["kspace"]
        if f.ncomp > 1:
            f.div_free()
    ti = integ(P, CFL=0.3)
    ti.save_cadence, ti.max_save_period, ti.iteration = 10 ** 9, 1e300, 1
    step = (lambda: ti.advance(data)) if lazy_dt else (lambda: ti.do_advance(data, 1e-4))
    for _ in range(3):
        step()
    calls = {}
    saved = (ts.lib, ph.lib)
    ts.lib = ph.lib = Stub(L.lib, calls)
    with np.errstate(divide="ignore"):
        t0 = time.perf_counter()
        for _ in range(n):
            step()
        us = (time.perf_counter() - t0) / n * 1e6
    ts.lib, ph.lib = saved
    print("%-20s %-13s %-7s %-18s %6.1f us per step   library calls per step: %s" % (
        physics.__name__, "x".join(map(str, shape)), integ.__name__, "advance(dt=None)" if lazy_dt else "do_advance(dt)", us,
        {k: round(v / n, 2) for k, v in calls.items()}))


if __name__ == "__main__":
    run(IncompressibleHydro, (128, 128), RK2mid)
    run(IncompressibleMHD, (512, 512), RK4, n=500)
    run(IncompressibleMHD, (64, 64, 64), RK4, n=500)
    run(BoussinesqHydro, (32, 32, 32), RK4, n=500)
    run(IncompressibleHydro, (128, 128), RK2mid, lazy_dt=True, n=1000)
