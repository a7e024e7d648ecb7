#!/usr/bin/env python
"""Goldens for the passive tracer carried by BoussinesqHydro / IncompressibleMHD, produced by RUNNING THE REFERENCE
(oracle/_ref): in the reference the tracer lives in IncompressibleHydro and is inherited (physics.py:467-470 field 'c'
inserted after 'u', :515-522 its diffusion, :578-583 its RHS inside the inherited hydro RHS).
Run in the build container only:   python tests/golden/make_tracer_goldens.py  ->  tests/golden/samples/options_tracer.npz"""
import os

import numpy as np

import make_sample_goldens as g

out = {}
g.decfg.set('physics', 'use_tracer', 'True')
try:
    for tag, name, shape, params, integ, seed in [
            ("tracer_bouss3d", "BoussinesqHydro", (16, 16, 16), dict(nu=1e-2, kappa=2e-2, c_diff=5e-3), "RK2mid", 41),
            ("tracer_mhd3d", "IncompressibleMHD", (16, 16, 16), dict(nu=1e-2, eta=2e-2, c_diff=1e-2), "RK2trap", 42),
            ("tracer_mhd2d", "IncompressibleMHD", (16, 32), dict(nu=1e-2, eta=1e-2, c_diff=0.), "RK2mid", 43)]:
        RHS, d = g.physics(name, shape, params)
        g.noise_state(d, seed)
        y0 = g.kvec(d)
        ti = getattr(g.ts, integ)(RHS)
        for _ in range(3):
            ti.do_advance(d, 1e-2)
        out[tag + "_y0"], out[tag + "_y1"] = y0, g.kvec(d)
        out[tag + "_fields"] = np.array([fn for fn, f in d])
        print("%-16s fields %s |y1| = %.12e" % (tag, [fn for fn, f in d], np.linalg.norm(out[tag + "_y1"])))
finally:
    g.decfg.set('physics', 'use_tracer', 'False')
np.savez_compressed(os.path.join(g.OUT, "options_tracer.npz"), **out)
