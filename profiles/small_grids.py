#!/usr/bin/env python
"""BASELINE configs 1-2 (2-D, L2-resident): microseconds per step, eager launches vs CUDA-graph replay."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dedalus-1.0_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import dedalus_oracle as orc
from devutil import dev_physics, oracle_physics, set_state
import dedalus.time_stepping.api as tapi

for name, physics, shape, integ, params in (("config 1: TG hydro 128^2 RK2mid", "IncompressibleHydro", (128, 128), "RK2mid", dict(nu=1e-3)),
                                            ("config 2: OT MHD 512^2 RK4", "IncompressibleMHD", (512, 512), "RK4", dict(nu=1e-3, eta=1e-3))):
    Po = oracle_physics(physics, shape, None, params)
    y0 = (orc.orszag_tang(Po.create_fields(0.)) if "MHD" in physics else orc.synthetic_ic(Po, 1)).kvector()
    out = {"config": name}
    for mode in ("eager", "graph"):
        P = dev_physics(physics, shape, None, params)
        data = P.create_fields(0.)
        set_state(data, y0)
        ti = getattr(tapi, integ)(P)
        step = ti.do_advance_graph if mode == "graph" else ti.do_advance
        for _ in range(5):
            step(data, 1e-3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 200
        e0.record()
        for _ in range(n):
            step(data, 1e-3)
        e1.record()
        torch.cuda.synchronize()
        out[mode + "_us_per_step"] = round(e0.elapsed_time(e1) / n * 1e3, 1)
    nk = (shape[1] // 2 + 1) * shape[0]
    stages = 4 if integ == "RK4" else 2
    out["graph_mode_stage_updates_per_s"] = stages * nk / (out["graph_us_per_step"] * 1e-6)
    print(json.dumps(out))
