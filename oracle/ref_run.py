#!/usr/bin/env python
"""Advance a given spectral state with the REFERENCE's own code and dump what it produced.

TEST INFRASTRUCTURE ONLY (executed by tests/ as a child process, because oracle/_ref's package
is also called `dedalus`).  Nothing on the product path may import oracle/.

What runs is the reference itself (oracle/_ref: its physics.py RHS, its representations.py with
the numpy FFT backend, its verbatim-compiled Cython stage kernels):

  * RK2mid / RK2trap: the reference's integrator classes, unmodified (time_step.py:275-392);
  * CrankNicholsonVisc: likewise restated (time_step.py:486-506) around the reference's RHS (ref_bench.restated_cn);
  * RK4: the reference cannot run its own (time_step.py:209,214,449), so the
    data flow of time_step.py:426-483 is restated around the reference's RHS and its Cython
    euler / etd1 kernels (oracle/ref_bench.py restated_rk4).

Input : --y0 FILE.npy   complex128 [n_components][k-space shape] in StateData insertion order
Output: --out DIR       y1.npy (same layout), meta.json (time, ekin, emag, divergence_sum, mag_div_sum
                        by the reference's own analysis/volume_average.py tasks, wall seconds)
`--threads T` serves the reference's numpy.fft calls by scipy.fft with T workers (same pocketfft
algorithm, the reference's code untouched), so that 256^3 fits a test's time budget.
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
import ref_bench  # noqa: E402


def run(a):
    decfg, data_api, physics_api, ts = build_ref.import_ref()
    from dedalus.data_objects.api import FourierRepresentation
    import dedalus.data_objects.representations as rep_mod
    import dedalus.analysis.volume_average as va
    if a.threads > 1:
        rep_mod.npfft = ref_bench._ThreadedFFT(a.threads)
    if a.direction:
        decfg.set("physics", "boussinesq_direction", a.direction)
    shape = tuple(a.shape)
    RHS = getattr(physics_api, a.physics)(shape, FourierRepresentation, tuple(a.length) if a.length else None)
    for kv in a.param:
        k, v = kv.split("=")
        RHS.parameters[k] = float(v)
    data = RHS.create_fields(0.)
    y0 = np.load(a.y0, mmap_mode="r")
    comps = [c for fn, f in data for i, c in f]
    assert y0.shape[0] == len(comps), (y0.shape, len(comps))
    for j, c in enumerate(comps):
        c['kspace'] = np.array(y0[j])
    if a.integ == "RK4":
        ti = ref_bench.restated_rk4(ts, RHS)
    elif a.integ == "CrankNicholsonVisc":
        ti = ref_bench.restated_cn(ts, RHS)
    else:
        ti = getattr(ts, a.integ)(RHS)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ti.do_advance(data, a.dt)
    sec = time.perf_counter() - t0
    os.makedirs(a.out, exist_ok=True)
    y1 = np.stack([np.array(c['kspace']) for c in comps])
    np.save(os.path.join(a.out, "y1.npy"), y1)
    # the reference's own diagnostics (register_task returns None, volume_average.py:68-69: tasks live in known_analysis)
    scratch = data.clone()
    scratch.add_field('scalar', 'ScalarField')
    task = va.VolumeAverageSet.known_analysis
    meta = {"time": float(data.time), "seconds": sec, "steps": a.steps, "threads": a.threads,
            "ekin": float(np.real(task["ekin"](data, scratch))),
            "divergence_sum": float(np.real(task["divergence_sum"](data, scratch)))}
    if a.physics == "IncompressibleMHD":
        meta["emag"] = float(np.real(task["emag"](data, scratch)))
        meta["mag_div_sum"] = float(np.real(task["mag_div_sum"](data, scratch)))
    with open(os.path.join(a.out, "meta.json"), "w") as f:
        json.dump(meta, f)
    print(json.dumps(meta))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--physics", default="IncompressibleMHD")
    ap.add_argument("--shape", type=int, nargs="+", required=True)
    ap.add_argument("--length", type=float, nargs="*", default=None)
    ap.add_argument("--integ", default="RK4")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--dt", type=float, required=True)
    ap.add_argument("--param", action="append", default=[], help="name=value into RHS.parameters")
    ap.add_argument("--direction", default=None, help="boussinesq_direction")
    ap.add_argument("--threads", type=int, default=1)
    ap.add_argument("--y0", required=True)
    ap.add_argument("--out", required=True)
    run(ap.parse_args())
