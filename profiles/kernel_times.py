#!/usr/bin/env python
"""Per-kernel CUDA-event times of the MHD RK4 step for a list of library option settings.
   python profiles/kernel_times.py --n 512 --opt xfused_variant=0,1,2
Used to choose launch variants on the B200 (numbers recorded in profiles/*.md)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dedalus-1.0_b200"))

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--opt", default="")
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()

import torch
import bench
import dedalus._lib as L
from dedalus.mods import RK4

P, data, dt = bench.make_state(a.n)
ti = RK4(P)
name, vals = (a.opt.split("=") + [""])[:2] if a.opt else ("", "")
for v in (vals.split(",") if vals else [None]):
    if v is not None:
        L.set_option(name, int(v))
    for _ in range(2):
        ti.do_advance(data, dt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        ti.do_advance(data, dt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    L.profile(True)
    ti.do_advance(data, dt)
    prof = L.profile_report()
    L.profile(False)
    print(json.dumps({"opt": "%s=%s" % (name, v), "ms_per_step": round(ms, 3),
                      "kernels_ms": {k: round(x["ms"] / x["n"], 3) for k, x in prof.items()}}))
