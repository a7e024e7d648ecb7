#!/usr/bin/env python
"""Generate tests/golden/samples/*.npz by RUNNING THE REFERENCE ITSELF (oracle/_ref): its initial-condition
generators (dedalus/init_cond/init_cond.py), its CFL-controlled `advance(data)` loop
(time_step.py:100-107,170-179 -> physics.py:151-158,601-610,714-721,821-836 -> fields.py:153-157) and its
VolumeAverageSet tasks (analysis/volume_average.py) -- the pieces the reference's sample scripts
(samples/incompressible_hydro/2d_decaying_turbulence, samples/incompressible_mhd/alfven_wave,
samples/boussinesq_hydro/gravity_wave) are made of, at sizes the CPU finishes in seconds.

Run in the build container only:   python tests/golden/make_sample_goldens.py
Random inputs: the reference draws from numpy's global generator (init_cond.py:312-322,455); every case
seeds it (`np.random.seed`) first, and the drop-in package draws the same numbers the same way.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "samples")
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import build_ref  # noqa: E402

decfg, data_api, physics_api, ts = build_ref.import_ref()
from dedalus.data_objects.api import FourierRepresentation  # noqa: E402
import dedalus.analysis.volume_average as va  # noqa: E402
import dedalus.init_cond.init_cond as ic  # noqa: E402
from dedalus.init_cond.turb_spectra import mcwilliams_spec  # noqa: E402


def kvec(data):
    return np.stack([c['kspace'].copy() for fn, f in data for i, c in f])


def physics(name, shape, params=None, direction=None):
    decfg.set('FFT', 'dealiasing', '2/3 cython')
    decfg.set('physics', 'boussinesq_direction', direction or ('y' if len(shape) == 2 else 'z'))
    RHS = getattr(physics_api, name)(shape, FourierRepresentation)
    RHS.parameters.update(params or {})
    return RHS, RHS.create_fields(0.)


def tasks(data, names):
    scratch = data.clone()
    scratch.add_field('scalar', 'ScalarField')
    known = va.VolumeAverageSet.known_analysis
    out = []
    for n in names:
        v = known[n](data, scratch)
        out.append(complex(v))
    return np.array(out)


def ic_cases():
    out = {}
    RHS, d = physics("IncompressibleHydro", (16, 16, 16))
    ic.taylor_green(d)
    out["tg3d"] = kvec(d)

    # alfven() cannot run in the reference as shipped: its diagnostic print reads a non-existent attribute
    # (init_cond.py:193-194, `data['u']['x'].shape`) before the perturbation is written -> no golden;
    # the port is pinned by the Alfven period instead (tests/test_gpu_widen.py)

    np.random.seed(1234)
    RHS, d = physics("IncompressibleHydro", (32, 32))
    ic.turb_new(d, mcwilliams_spec, k0=5., E0=1.)
    out["turb2d"] = kvec(d)

    np.random.seed(4321)
    RHS, d = physics("IncompressibleHydro", (16, 16, 32))
    ic.turb_new(d, mcwilliams_spec, tot_en=0.7, k0=4., E0=1.)
    out["turb3d"] = kvec(d)

    RHS, d = physics("IncompressibleHydro", (32, 32))
    ic.MIT_vortices(d)
    out["mit"] = kvec(d)

    # add_gaussian_white_noise() ends in enforce_hermitian(), which gathers through mpi4py unconditionally
    # (representations.py:480): not runnable without MPI -> no golden; the port is pinned by properties
    # (realness of the x-space field, <noise^2> = std^2) in tests/test_gpu_widen.py

    # the older 2-D generator turb() followed by remove_compressible() (init_cond.py:343-389)
    np.random.seed(99)
    RHS, d = physics("IncompressibleHydro", (24, 32))
    ic.turb(d['u']['x'], d['u']['y'], mcwilliams_spec, k0=4., E0=1.)
    out["turb_old"] = kvec(d)
    ic.remove_compressible(d['u']['x'], d['u']['y'])
    out["turb_old_projected"] = kvec(d)

    RHS, d = physics("BoussinesqHydro", (16, 16))
    ic.sin_k(d['u']['x']['kspace'], (2, 3), ampl=0.5)
    ic.cos_k(d['u']['y']['kspace'], (1, 2), ampl=-1.5)
    ic.constant(d, 'T', 2.5)
    ic.constant(d, 'uy', -0.75)
    out["sincos"] = kvec(d)
    np.savez_compressed(os.path.join(OUT, "init_cond.npz"), **out)
    print("init_cond.npz:", {k: float(np.linalg.norm(v)) for k, v in out.items()})


RUNS = [
    # tag, physics, shape, params, integrator, CFL, nsteps, ic, task names
    ("cfl_turb2d", "IncompressibleHydro", (32, 32), dict(nu=1e-3), "RK2mid", 0.4, 5, "turb2d",
     ["ekin", "enstrophy", "vort_cenk", "ux2", "uy2", "divergence", "divergence_sum"]),
    ("cfl_mhd3d", "IncompressibleMHD", (16, 16, 16), dict(nu=1e-2, eta=2e-2, rho0=0.8), "RK2mid", 0.3, 4, "mhd3d",
     ["ekin", "emag", "ux2", "uy2", "uz2", "bx2", "by2", "bz2", "divergence_sum", "mag_div_sum", "energy_dissipation",
      "divergence", "mag_div"]),
    ("cfl_bouss2d", "BoussinesqHydro", (16, 32), dict(nu=1e-3, kappa=2e-3, g=2.0, beta=3.0), "RK2trap", 0.5, 4, "gmode",
     ["ekin", "temp2", "ux2", "uy2", "divergence_sum"]),
    ("cfl_bouss3d", "BoussinesqHydro", (16, 16, 16), dict(nu=1e-2, kappa=1e-2), "RK2mid", 0.35, 3, "bouss3d",
     ["ekin", "temp2", "thermal_energy_dissipation", "energy_dissipation"]),
]


def set_ic(kind, RHS, data):
    if kind == "turb2d":
        np.random.seed(1234)
        ic.turb_new(data, mcwilliams_spec, k0=5., E0=1.)
    elif kind == "mhd3d":
        np.random.seed(77)
        ic.turb_new(data, mcwilliams_spec, tot_en=0.5, k0=3., E0=1.)
        rng = np.random.default_rng(5)
        for i, c in data['B']:
            c['xspace'] = rng.standard_normal(tuple(c.local_shape['xspace']))
            c['kspace']
        data['B'].div_free()
    elif kind == "gmode":
        # samples/boussinesq_hydro/gravity_wave/2d_gmode_kx1_kz1.py: one (kx, ky) = (1, 1) mode in T and u
        ic.sin_k(data['T']['kspace'], (1, 1), ampl=0.1)
        ic.cos_k(data['u']['x']['kspace'], (1, 1), ampl=0.05)
        ic.cos_k(data['u']['y']['kspace'], (1, 1), ampl=-0.05)
    elif kind == "bouss3d":
        rng = np.random.default_rng(8)
        for fn, f in data:
            for i, c in f:
                c['xspace'] = rng.standard_normal(tuple(c.local_shape['xspace']))
                c['kspace']
        data['u'].div_free()


def run_cases():
    for tag, phys, shape, params, integ, cfl, nsteps, kind, names in RUNS:
        RHS, data = physics(phys, shape, params)
        set_ic(kind, RHS, data)
        y0 = kvec(data)
        t0 = tasks(data, names)
        ti = getattr(ts, integ)(RHS, CFL=cfl)
        ti.save_cadence, ti.max_save_period, ti.iteration = 10 ** 9, 1e300, 1      # no snapshots
        dts = []
        for _ in range(nsteps):
            ti.advance(data)               # dt = None: cfl_dt -> compute_dt -> max_square
            dts.append(ti.dt_old)
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), y0=y0, y1=kvec(data), dts=np.array(dts), time=data.time,
                            tasks0=t0, tasks1=tasks(data, names), task_names=np.array(names),
                            meta=np.array(repr(dict(physics=phys, shape=shape, params=params, integ=integ, CFL=cfl,
                                                    nsteps=nsteps))))
        print("%-12s dts=%s t=%.6g |y1|=%.12e" % (tag, ["%.5e" % d for d in dts], data.time, np.linalg.norm(kvec(data))))


def noise_state(data, seed, solenoidal=True):
    rng = np.random.default_rng(seed)
    for fn, f in data:
        for i, c in f:
            c['xspace'] = 0.5 * rng.standard_normal(tuple(c.local_shape['xspace']))
            c['kspace']
        if solenoidal and f.ncomp > 1:
            f.div_free()


def forcing_arrays(RHS, seed, n):
    """n fixed, dealiased, Hermitian-consistent spectra (the forward transform of real noise)."""
    aux = RHS.create_fields(0.)
    rng = np.random.default_rng(seed)
    out = []
    c = aux['u'][0]
    for _ in range(n):
        c['xspace'] = 0.2 * rng.standard_normal(tuple(c.local_shape['xspace']))
        out.append(c['kspace'].copy())
    return out


def option_cases():
    """Rotation, forcing and passive-tracer branches of the reference RHS (physics.py:560-586,709-711), fixed dt."""
    out = {}

    def run(tag, RHS, data, integ, dt, nsteps, extra=None):
        y0 = kvec(data)
        ti = getattr(ts, integ)(RHS)
        for _ in range(nsteps):
            ti.do_advance(data, dt)
        out[tag + "_y0"], out[tag + "_y1"] = y0, kvec(data)
        for k, v in (extra or {}).items():
            out[tag + "_" + k] = v
        print("%-10s |y1|=%.12e" % (tag, np.linalg.norm(out[tag + "_y1"])))

    RHS, d = physics("IncompressibleHydro", (16, 16, 16), dict(nu=1e-2, Omega=np.array([1.0, 0.2, 0.3])))
    noise_state(d, 31)
    run("rot3d", RHS, d, "RK2mid", 1e-2, 3)

    RHS, d = physics("IncompressibleHydro", (16, 32), dict(nu=1e-2, Omega=0.7))
    noise_state(d, 32)
    run("rot2d", RHS, d, "RK2trap", 1e-2, 3)

    RHS, d = physics("IncompressibleHydro", (16, 32), dict(nu=1e-2))
    noise_state(d, 33)
    F = forcing_arrays(RHS, 34, 2)
    RHS.set_velocity_forcing(lambda data, i: F[i])
    run("force2d", RHS, d, "RK2mid", 1e-2, 3, dict(F=np.stack(F)))

    RHS, d = physics("BoussinesqHydro", (16, 16, 16), dict(nu=1e-2, kappa=1e-2))
    noise_state(d, 35)
    F = forcing_arrays(RHS, 36, 1)
    RHS.set_thermal_forcing(lambda data: F[0])
    run("heat3d", RHS, d, "RK2mid", 1e-2, 3, dict(F=np.stack(F)))

    decfg.set('physics', 'use_tracer', 'True')
    try:
        RHS, d = physics("IncompressibleHydro", (16, 32), dict(nu=1e-2, c_diff=2e-2))
        noise_state(d, 37)
        run("tracer2d", RHS, d, "RK2mid", 1e-2, 3)
        RHS, d = physics("IncompressibleHydro", (16, 16, 16), dict(nu=1e-2, c_diff=0.))
        noise_state(d, 38)
        run("tracer3d", RHS, d, "RK2trap", 1e-2, 3)
    finally:
        decfg.set('physics', 'use_tracer', 'False')
    np.savez_compressed(os.path.join(OUT, "options.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ic_cases()
    run_cases()
    option_cases()
