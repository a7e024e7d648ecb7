#!/usr/bin/env python
"""BASELINE configs 3-4 on one GPU (the bench line is config 5's physics at 512^3): ms/step and
mode-stage updates/s for 3-D hydro 256^3 RK4 and 3-D Boussinesq 512^3 RK4 / RK2mid, synthetic state."""
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
import dedalus.analysis.volume_average as va
import dedalus._lib as L

SHRINK = int(sys.argv[sys.argv.index("--shrink") + 1]) if "--shrink" in sys.argv else 1     # dry runs (tests/bench_emul_child.py)
A_STAGE = {"IncompressibleHydro": 1104.0, "BoussinesqHydro": 1568.0, "IncompressibleMHD": 1920.0}
for physics, n, integ, params in (("IncompressibleHydro", 256, "RK4", dict(nu=1e-3)),
                                  ("BoussinesqHydro", 512, "RK4", dict(nu=1e-3, kappa=1e-3)),
                                  ("BoussinesqHydro", 512, "RK2mid", dict(nu=1e-3, kappa=1e-3)),
                                  ("IncompressibleHydro", 512, "RK4", dict(nu=1e-3))):
    n //= SHRINK
    P = dev_physics(physics, (n, n, n), None, params)
    data = P.create_fields(0.)
    g = torch.Generator(device="cuda").manual_seed(3)
    kk = torch.sqrt(data["u"][0].k2())
    amp = torch.where(kk > 0, kk.clamp(min=1e-30) ** (-5.0 / 6.0), torch.zeros_like(kk))
    del kk
    for _, f in data:
        for _, c in f:
            c["xspace"] = torch.randn(n, n, n, dtype=torch.float64, device="cuda", generator=g)
            c["kspace"].mul_(amp)
            c._xdata = None
        if f.ncomp > 1:
            f.div_free()
        en = sum(va.volume_average(c["kspace"].abs() ** 2, kdict=c.k) for _, c in f)
        for _, c in f:
            c["kspace"].mul_(1.0 / np.sqrt(en))
    del amp
    umax = float(data["u"].max_square()) ** 0.5
    for _, c in data["u"]:
        c["kspace"]
        c._xdata = None
    torch.cuda.empty_cache()
    dt = 0.2 * (2 * np.pi / n) / umax
    ti = getattr(tapi, integ)(P)
    for _ in range(3):
        ti.do_advance(data, dt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        ti.do_advance(data, dt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    stages = 4 if integ == "RK4" else 2
    nk = n * n * (n // 2 + 1)
    ups = stages * nk / (ms * 1e-3)
    L.profile(True)
    ti.do_advance(data, dt)
    prof = L.profile_report()
    L.profile(False)
    print(json.dumps({"physics": physics, "grid": n, "integrator": integ, "ms_per_step": round(ms, 3), "updates_per_s": ups,
                      "algorithmic_tbs": round(A_STAGE[physics] * ups / 1e12, 2),
                      "kernels_ms": {k: round(v["ms"] / v["n"], 3) for k, v in prof.items()}}))
    del data, ti, P
    torch.cuda.empty_cache()
