#!/usr/bin/env python
"""ms/step of the rows widened at the end of round 1, for the first GPU call of round 2 (nothing here has been timed yet):
grids that are not powers of two on the runtime-length kernels (the reference's sample grids 450^2 and 48 x 2 x 48, then 384^3 =
2^7 * 3 against its power-of-two neighbours) and the shearing box (FourierShearRepresentation: unfused helper sequence + tensor-level
ETD stages, RK2mid).  One JSON line per case;  python profiles/widened_configs.py [--quick]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dedalus-1.0_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
from devutil import dev_physics
import dedalus.time_stepping.api as tapi
import dedalus._lib as L
from dedalus.data_objects.api import FourierShearRepresentation

quick = "--quick" in sys.argv
CASES = [  # physics, shape, integrator, params, shear rate (None: static representation)
    ("IncompressibleHydro", (450, 450), "RK2mid", dict(nu=1e-4), None),
    ("BoussinesqHydro", (48, 2, 48), "RK2mid", dict(nu=1e-3, kappa=1e-3), None),
    ("IncompressibleMHD", (256, 256, 256), "RK4", dict(nu=1e-3, eta=1e-3), None),
    ("IncompressibleMHD", (384, 384, 384), "RK4", dict(nu=1e-3, eta=1e-3), None),
    ("IncompressibleHydro", (512, 512), "RK2mid", dict(nu=1e-4), 1.5),
    ("IncompressibleHydro", (512, 512), "RK2mid", dict(nu=1e-4), None),
    ("IncompressibleMHD", (128, 128, 128), "RK2mid", dict(nu=1e-3, eta=1e-3), 1.5),
    ("IncompressibleMHD", (128, 128, 128), "RK2mid", dict(nu=1e-3, eta=1e-3), None),
]
if quick:
    CASES = [c for c in CASES if np.prod(c[1]) <= 512 * 512]

for physics, shape, integ, params, S in CASES:
    if S is None:
        P = dev_physics(physics, shape, None, params)
    else:
        import dedalus.physics.api as papi
        from dedalus.config import decfg
        decfg.set("FFT", "dealiasing", "2/3 cython")
        P = getattr(papi, physics)(tuple(shape), FourierShearRepresentation)
        P.parameters.update(params)
        P.parameters["shear_rate"] = S
    data = P.create_fields(0.)
    g = torch.Generator(device="cuda").manual_seed(3)
    for _, f in data:
        for _, c in f:
            c["xspace"] = torch.randn(*shape, dtype=torch.float64, device="cuda", generator=g)
            c["kspace"]
        if f.ncomp > 1:
            f.div_free()
    umax = float(data["u"].max_square()) ** 0.5
    for _, _, c in data.components():
        c["kspace"]
    dt = 0.2 * (2 * np.pi / max(shape)) / umax
    ti = getattr(tapi, integ)(P)
    for _ in range(3):
        ti.do_advance(data, dt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    l0 = L.launch_count()
    e0.record()
    for _ in range(steps):
        ti.do_advance(data, dt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    nk = int(np.prod(shape[:-1])) * (shape[-1] // 2 + 1)
    stages = 4 if integ == "RK4" else 2
    print(json.dumps({"physics": physics, "shape": list(shape), "integrator": integ, "shear_rate": S, "ms_per_step": round(ms, 3),
                      "updates_per_s": stages * nk / (ms * 1e-3), "library_launches_per_step": (L.launch_count() - l0) / steps}), flush=True)
    del data, ti, P
    torch.cuda.empty_cache()
