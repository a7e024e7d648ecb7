#!/usr/bin/env python
"""Time the reference's own CPU implementation of the hot path (oracle/_ref: the reference's
Python + verbatim-compiled Cython kernels, numpy-FFT backend) on this host's cores.

TEST / BENCH INFRASTRUCTURE: executed only by bench.py (`--impl reference` and the
`cpu_baseline` leg).  Runs in its own process because oracle/_ref's package is also called
`dedalus`.

RK4 does not run in the reference as shipped (time_step.py:209,214,449); the
"RK4" timed here is the restated glue of SURVEY.md section 8(c) around the REFERENCE's
RHS (physics.py) and the REFERENCE's Cython euler/etd1 kernels -- labelled "restated".
The path is single-threaded by construction (numpy pocketfft; the reference has no threading);
its MPI / FFTW-MPI mode cannot be reproduced in this image (no MPI, no FFTW).  `--threads T` is the
best-effort use of the host's cores without touching the reference's code: its numpy.fft calls are served by
scipy.fft with T workers (same pocketfft algorithm); everything else stays as it is.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_ref  # noqa: E402


def restated_rk4(ts_mod, RHS):
    """RK4 data flow of time_step.py:426-483 with distinct k buffers; forward_step
    (:187-221) with linear_step = euler, intfac_step = etd1(-IF)."""
    import forward_step_cy_2d as f2
    import forward_step_cy_3d as f3
    mod = f2 if RHS.ndim == 2 else f3

    class RK4(ts_mod.TimeStepBase):
        def __init__(self, RHS):
            ts_mod.TimeStepBase.__init__(self, RHS)
            self.tmp, self.k, self.tot = (RHS.create_fields(0.) for _ in range(3))

        def _step(self, start, deriv, out, dt):
            for fn, f in out:
                for i, c in f:
                    IF = self.k[fn][i].integrating_factor
                    if IF is None:
                        mod.euler(start[fn][i]['kspace'], c['kspace'], deriv[fn][i]['kspace'], dt)
                    else:
                        mod.etd1(start[fn][i]['kspace'], c['kspace'], deriv[fn][i]['kspace'], -IF, dt)
            out.set_time(start.time + dt)

        def do_advance(self, data, dt):
            R, k, tot, tmp = self.RHS, self.k, self.tot, self.tmp
            for w, h, first in ((6., dt / 2., True), (3., dt / 2., False), (3., dt, False), (6., None, False)):
                R.RHS(data if first else tmp, k)
                for fn, f in tot:
                    for i, c in f:
                        if first:
                            c['kspace'] = k[fn][i]['kspace'] / w
                        else:
                            c['kspace'] += k[fn][i]['kspace'] / w
                if h is not None:
                    self._step(data, k, tmp, h)
            self._step(data, tot, data, dt)
            self.time += dt
            self.iteration += 1
    return RK4(RHS)


def restated_cn(ts_mod, RHS):
    """CrankNicholsonVisc.do_advance (time_step.py:486-506, which cannot run as shipped) restated around the
    REFERENCE's RHS: per component y+ = ((1/dt - IF/2) y + N(y)) / (1/dt + IF/2), IF = 0 where the factor is None."""

    class CN(ts_mod.TimeStepBase):
        def __init__(self, RHS):
            ts_mod.TimeStepBase.__init__(self, RHS)
            self.deriv = RHS.create_fields(0.)

        def do_advance(self, data, dt):
            self.RHS.RHS(data, self.deriv)
            for fn, f in data:
                for i, c in f:
                    IF = self.deriv[fn][i].integrating_factor
                    IF = 0. if IF is None else IF
                    c['kspace'] = ((1. / dt - IF / 2.) * c['kspace'] + self.deriv[fn][i]['kspace']) / (1. / dt + IF / 2.)
            data.set_time(data.time + dt)
            self.time += dt
            self.iteration += 1
    return CN(RHS)


class _ThreadedFFT(object):
    """numpy.fft's interface on scipy.fft with `workers` threads: the one part of the reference's numpy route
    (representations.py:327-333) that can use more than one core without touching its code."""

    def __init__(self, workers):
        import scipy.fft as sfft
        self._s, self._w = sfft, workers

    def __getattr__(self, name):
        f = getattr(self._s, name)
        if name in ("fftfreq", "rfftfreq", "fftshift", "ifftshift"):
            return f
        return lambda *a, **kw: f(*a, workers=self._w, **kw)


def run(n, ndim, steps, warmup, physics="IncompressibleMHD", integ="RK4", threads=1):
    decfg, data_api, physics_api, ts = build_ref.import_ref()
    from dedalus.data_objects.api import FourierRepresentation
    import dedalus.data_objects.representations as rep_mod
    if threads > 1:
        rep_mod.npfft = _ThreadedFFT(threads)
    shape = (n,) * ndim
    RHS = getattr(physics_api, physics)(shape, FourierRepresentation)
    RHS.parameters['nu'] = 1e-3
    if physics == "IncompressibleMHD":
        RHS.parameters['eta'] = 1e-3
    data = RHS.create_fields(0.)
    idx = 0
    for fn, f in data:
        for i, c in f:
            rng = np.random.default_rng(5000 + idx)
            c['xspace'] = rng.standard_normal(shape)
            c['kspace']
            idx += 1
        if f.ncomp > 1:
            f.div_free()
    umax = max(np.abs(c['xspace']).max() for i, c in data['u'])
    for fn, f in data:
        for i, c in f:
            c['kspace']
    dt = 0.2 * (2 * np.pi / n) / umax
    ti = restated_rk4(ts, RHS) if integ == "RK4" else getattr(ts, integ)(RHS)
    nstage = 4 if integ == "RK4" else 2
    for _ in range(warmup):
        ti.do_advance(data, dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        ti.do_advance(data, dt)
    sec = time.perf_counter() - t0
    nk = (n // 2 + 1) * n ** (ndim - 1)
    assert np.isfinite(data['u'][0]['kspace']).all()
    return dict(value=steps * nstage * nk / sec, ms_per_step=1e3 * sec / steps, n=n, ndim=ndim, steps=steps, threads=threads,
                warmup=warmup, nk=nk, integrator=integ + (" (restated)" if integ == "RK4" else ""), physics=physics)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--ndim", type=int, default=3)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=1, help="FFT worker threads (scipy.fft behind the reference's numpy.fft calls)")
    a = ap.parse_args()
    print(json.dumps(run(a.n, a.ndim, a.steps, a.warmup, threads=a.threads)))
