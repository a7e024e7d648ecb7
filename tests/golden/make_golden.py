#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (oracle/_ref, i.e. the
reference's own Python + verbatim-compiled Cython kernels, numpy-FFT backend).

Run in the build container only (needs /root/reference once, to materialise oracle/_ref):
    python tests/golden/make_golden.py
The .npz files are committed; the GPU box never needs /root/reference.

Each case stores the initial spectral state (all components, StateData insertion order),
the time derivative after ONE reference RHS call, the state after n steps of the named
reference integrator, and the reference's invariants (volume_average.py tasks).
RK4 / CrankNicholsonVisc are not runnable in the reference and therefore
have no golden vectors.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import build_ref  # noqa: E402

decfg, data_api, physics_api, ts = build_ref.import_ref()
from dedalus.data_objects.api import FourierRepresentation  # noqa: E402
import dedalus.analysis.volume_average as va  # noqa: E402
import dedalus.init_cond.init_cond as ic  # noqa: E402
import forward_step_cy_2d as fs2  # noqa: E402
import forward_step_cy_3d as fs3  # noqa: E402


CASES = [
    # name, physics, shape (z,y,x)/(y,x), length, params, integrator, dt, nsteps, ic, dealiasing
    dict(name="hydro2d_16_rk2mid", physics="IncompressibleHydro", shape=(16, 16), length=None,
         params=dict(nu=0.05), integ="RK2mid", dt=2e-2, nsteps=3, ic="synthetic", cfg=1),
    dict(name="tg2d_32_rk2mid", physics="IncompressibleHydro", shape=(32, 32), length=None,
         params=dict(nu=0.1), integ="RK2mid", dt=1e-2, nsteps=5, ic="taylor_green", cfg=1),
    dict(name="mhd2d_16x32_rk2mid", physics="IncompressibleMHD", shape=(16, 32), length=None,
         params=dict(nu=0.02, eta=0.03), integ="RK2mid", dt=1e-2, nsteps=3, ic="synthetic", cfg=2),
    dict(name="mhd2d_32_rk2trap_inviscid", physics="IncompressibleMHD", shape=(32, 32), length=None,
         params=dict(), integ="RK2trap", dt=5e-3, nsteps=3, ic="orszag_tang", cfg=2),
    dict(name="hydro3d_16_rk2mid", physics="IncompressibleHydro", shape=(16, 16, 16), length=None,
         params=dict(nu=0.01), integ="RK2mid", dt=1e-2, nsteps=3, ic="synthetic", cfg=3),
    dict(name="hydro3d_16_rk2trap", physics="IncompressibleHydro", shape=(16, 16, 16), length=None,
         params=dict(nu=0.01), integ="RK2trap", dt=1e-2, nsteps=3, ic="synthetic", cfg=3),
    dict(name="bouss3d_16_rk2mid", physics="BoussinesqHydro", shape=(16, 16, 16), length=None,
         params=dict(nu=0.01, kappa=0.02, g=1.3, alpha_t=0.7, beta=1.1), integ="RK2mid", dt=1e-2, nsteps=3,
         ic="synthetic", cfg=4),
    dict(name="mhd3d_16_rk2mid", physics="IncompressibleMHD", shape=(16, 16, 16), length=None,
         params=dict(nu=0.01, eta=0.01), integ="RK2mid", dt=5e-3, nsteps=3, ic="synthetic", cfg=5),
    dict(name="mhd3d_16_rk2mid_stiff", physics="IncompressibleMHD", shape=(16, 16, 16), length=None,
         params=dict(nu=0.5, eta=0.8), integ="RK2mid", dt=1e-2, nsteps=3, ic="synthetic", cfg=5),
    dict(name="mhd3d_16_rk2mid_inviscid", physics="IncompressibleMHD", shape=(16, 16, 16), length=None,
         params=dict(), integ="RK2mid", dt=5e-3, nsteps=3, ic="synthetic", cfg=5),
    dict(name="mhd3d_8x16x32_rk2trap", physics="IncompressibleMHD", shape=(8, 16, 32),
         length=(2.0, 3.0, 5.0), params=dict(nu=0.02, eta=0.01, rho0=1.7), integ="RK2trap", dt=2e-3, nsteps=2,
         ic="synthetic", cfg=5),
    dict(name="hydro3d_16_rk2mid_nodealias", physics="IncompressibleHydro", shape=(16, 16, 16), length=None,
         params=dict(nu=0.01), integ="RK2mid", dt=5e-3, nsteps=2, ic="synthetic", cfg=3, dealiasing="None"),
    # grids that are not powers of two (the reference transforms any N through FFTW / numpy.fft; its samples run
    # 30 x 10, 48 x 2 x 48, 450 x 450, 100 x 100): mixed radix 2 / 3 / 5, odd lengths, a length-2 axis, a factor 7
    dict(name="hydro2d_10x30_rk2mid", physics="IncompressibleHydro", shape=(10, 30), length=(2 * np.pi, 6 * np.pi),
         params=dict(nu=0.02), integ="RK2mid", dt=1e-2, nsteps=3, ic="synthetic", cfg=1),
    dict(name="hydro2d_50_rk2trap", physics="IncompressibleHydro", shape=(50, 50), length=None,
         params=dict(nu=0.01), integ="RK2trap", dt=5e-3, nsteps=2, ic="synthetic", cfg=1),
    dict(name="mhd2d_18x24_rk2mid", physics="IncompressibleMHD", shape=(18, 24), length=None,
         params=dict(nu=0.02, eta=0.03), integ="RK2mid", dt=1e-2, nsteps=3, ic="synthetic", cfg=2),
    dict(name="bouss3d_48x2x48_rk2mid", physics="BoussinesqHydro", shape=(48, 2, 48), length=None,
         params=dict(nu=0.01, kappa=0.02, g=1.3, alpha_t=0.7, beta=1.1), integ="RK2mid", dt=1e-2, nsteps=2,
         ic="synthetic", cfg=4),
    dict(name="mhd3d_9x15x14_rk2trap", physics="IncompressibleMHD", shape=(9, 15, 14), length=(2.0, 3.0, 5.0),
         params=dict(nu=0.02, eta=0.01, rho0=1.7), integ="RK2trap", dt=2e-3, nsteps=2, ic="synthetic", cfg=5),
    dict(name="hydro3d_12x20x24_rk2mid", physics="IncompressibleHydro", shape=(12, 20, 24), length=None,
         params=dict(nu=0.01), integ="RK2mid", dt=1e-2, nsteps=2, ic="synthetic", cfg=3),
    # FFT.dealiasing = None (Nyquist planes zeroed only, aliased products): more physics classes, 2-D, a grid that is not a power of two
    dict(name="mhd3d_16_rk2trap_nodealias", physics="IncompressibleMHD", shape=(16, 16, 16), length=None,
         params=dict(nu=0.01, eta=0.02), integ="RK2trap", dt=5e-3, nsteps=2, ic="synthetic", cfg=5, dealiasing="None"),
    dict(name="bouss2d_12x20_rk2mid_nodealias", physics="BoussinesqHydro", shape=(12, 20), length=None,
         params=dict(nu=0.01, kappa=0.02), integ="RK2mid", dt=5e-3, nsteps=2, ic="synthetic", cfg=4, dealiasing="None"),
    # FFT.dealiasing = '2/3 spherical' (representations.py:410-417), a box that is not cubic so that min(kny) matters
    dict(name="mhd3d_16_rk2mid_nodealias_spherical", physics="IncompressibleMHD", shape=(16, 16, 16), length=(2 * np.pi, 3.0, 2 * np.pi),
         params=dict(nu=0.01, eta=0.02), integ="RK2mid", dt=5e-3, nsteps=2, ic="synthetic", cfg=5, dealiasing="2/3 spherical"),
    dict(name="hydro2d_24x16_rk2trap_nodealias_spherical", physics="IncompressibleHydro", shape=(24, 16), length=None,
         params=dict(nu=0.01), integ="RK2trap", dt=5e-3, nsteps=2, ic="synthetic", cfg=1, dealiasing="2/3 spherical"),
    dict(name="hydro2d_16_rk2mid_visc2", physics="IncompressibleHydro", shape=(16, 16), length=None,
         params=dict(nu=1e-3, viscosity_order=2), integ="RK2mid", dt=1e-2, nsteps=3, ic="synthetic", cfg=1),
]


def kvec(data):
    return np.stack([c['kspace'].copy() for fn, f in data for i, c in f])


def synthetic(RHS, data, cfg):
    """SURVEY.md 8(d) recipe executed through the reference's own objects."""
    c0 = data['u'][0]
    kk = np.sqrt(c0.k2())
    shape = np.zeros_like(kk)
    shape[kk > 0] = kk[kk > 0] ** (-5.0 / 6.0)
    idx = 0
    for fn, f in data:
        for i, c in f:
            rng = np.random.default_rng(1000 * cfg + idx)
            c['xspace'] = rng.standard_normal(tuple(c.local_shape['xspace']))
            c['kspace'] *= shape
            idx += 1
        if f.ncomp > 1:
            f.div_free()
        en = sum(va.volume_average(np.abs(c['kspace']) ** 2, kdict=c.k).real for i, c in f)
        for i, c in f:
            c['kspace'] *= 1.0 / np.sqrt(en)


def orszag_tang(data):
    c0 = data['u'][0]
    ny, nx = c0.local_shape['xspace']
    L = c0.length
    y = (np.arange(ny) * L[0] / ny)[:, None] * np.ones((1, nx))
    x = (np.arange(nx) * L[1] / nx)[None, :] * np.ones((ny, 1))
    data['u'][0]['xspace'] = -np.sin(y)
    data['u'][1]['xspace'] = np.sin(x)
    data['B'][0]['xspace'] = -np.sin(y)
    data['B'][1]['xspace'] = np.sin(2 * x)
    for fn, f in data:
        for i, c in f:
            c['kspace']


def invariants(data):
    scratch = data.clone()
    scratch.add_field('scalar', 'ScalarField')
    # register_task returns None (volume_average.py:68-69): tasks live in known_analysis
    task = va.VolumeAverageSet.known_analysis
    out = {"ekin": task["ekin"](data, scratch), "divergence_sum": task["divergence_sum"](data, scratch)}
    if 'B' in data.fields:
        out["emag"] = task["emag"](data, scratch)
        out["mag_div_sum"] = task["mag_div_sum"](data, scratch)
    return {k: float(np.real(v)) for k, v in out.items()}


def run_case(c):
    decfg.set('FFT', 'dealiasing', c.get("dealiasing", "2/3 cython"))
    if c["physics"] == "BoussinesqHydro" and len(c["shape"]) == 2:
        decfg.set('physics', 'boussinesq_direction', 'y')
    else:
        decfg.set('physics', 'boussinesq_direction', 'z')
    P = getattr(physics_api, c["physics"])
    RHS = P(c["shape"], FourierRepresentation, c["length"])
    RHS.parameters.update(c["params"])
    data = RHS.create_fields(0.)
    if c["ic"] == "synthetic":
        synthetic(RHS, data, c["cfg"])
    elif c["ic"] == "taylor_green":
        ic.taylor_green(data)
    elif c["ic"] == "orszag_tang":
        orszag_tang(data)
    y0 = kvec(data)
    inv0 = invariants(data)
    # one RHS on a copy (the reference's RHS mutates MHD state in place, F5)
    RHS2 = P(c["shape"], FourierRepresentation, c["length"])
    RHS2.parameters.update(c["params"])
    d2 = RHS2.create_fields(0.)
    for (fn, f), (_, g) in zip(data, d2):
        for (i, a), (_, b) in zip(f, g):
            b['kspace'] = a['kspace']
    deriv = RHS2.create_fields(0.)
    RHS2.RHS(d2, deriv)
    dy0 = kvec(deriv)
    y0_after_rhs = kvec(d2)
    ti = getattr(ts, c["integ"])(RHS)
    for _ in range(c["nsteps"]):
        ti.do_advance(data, c["dt"])
    y1 = kvec(data)
    inv1 = invariants(data)
    meta = dict(c)
    meta["length"] = list(c["length"]) if c["length"] else None
    np.savez_compressed(
        os.path.join(HERE, c["name"] + ".npz"),
        y0=y0, dy0=dy0, y0_after_rhs=y0_after_rhs, y1=y1, time=data.time,
        inv0=np.array([inv0[k] for k in sorted(inv0)]), inv1=np.array([inv1[k] for k in sorted(inv1)]),
        inv_names=np.array(sorted(inv0)), meta=np.array(repr(meta)))
    print("%-32s |y1|=%.12e  t=%g  %s" % (c["name"], np.linalg.norm(y1), data.time, inv1))


def stage_kernel_vectors():
    """Direct outputs of the reference's Cython euler/etd1/etd2rk1/etd2rk2 (2-D and 3-D) on
    random inputs whose Z = IF*dt straddles the ==0, |Z|<0.5 and exp branches."""
    rng = np.random.default_rng(7)
    out = {}
    for nd, mod, shp in ((2, fs2, (9, 16)), (3, fs3, (6, 5, 9))):
        cplx = lambda: (rng.standard_normal(shp) + 1j * rng.standard_normal(shp))
        start, d1, d2 = cplx(), cplx(), cplx()
        IF = -np.abs(rng.standard_normal(shp)) * 40.0
        IF.flat[::7] = 0.0
        IF.flat[1::5] *= 0.01
        dt = 0.02
        o = np.zeros(shp, dtype=np.complex128)
        pre = "k%dd_" % nd
        out[pre + "start"], out[pre + "d1"], out[pre + "d2"], out[pre + "IF"], out[pre + "dt"] = start, d1, d2, IF, dt
        mod.euler(start, o, d1, dt); out[pre + "euler"] = o.copy()
        mod.etd1(start, o, d1, IF, dt); out[pre + "etd1"] = o.copy()
        mod.etd2rk1(start, o, d1, d2, IF, dt); out[pre + "etd2rk1"] = o.copy()
        mod.etd2rk2(start, o, d1, d2, IF, dt); out[pre + "etd2rk2"] = o.copy()
    np.savez_compressed(os.path.join(HERE, "stage_kernels.npz"), **out)
    print("stage_kernels.npz written")


def transform_vectors():
    """forward()/backward() of the reference representation incl. layout F9 and dealias."""
    out = {}
    for name, shp, L, dl in (("t2d", (16, 32), (2 * np.pi, 2 * np.pi), "2/3 cython"),
                             ("t3d", (8, 16, 32), (2.0, 3.0, 5.0), "2/3 cython"),
                             ("t3dn", (16, 16, 16), (2 * np.pi,) * 3, "None")):
        decfg.set('FFT', 'dealiasing', dl)
        rep = FourierRepresentation(None, shp, L)
        rng = np.random.default_rng(11)
        x = rng.standard_normal(shp)
        rep['xspace'] = x
        k = rep['kspace'].copy()
        xb = rep['xspace'].copy()
        out[name + "_x"], out[name + "_k"], out[name + "_xb"] = x, k, xb
        out[name + "_dx"] = rep.deriv('x').copy() if False else None
        rep['kspace']
        out[name + "_derivx"] = rep.deriv('x').copy()
        out[name + "_derivy"] = rep.deriv('y').copy()
        out[name + "_k2"] = rep.k2()
        del out[name + "_dx"]
    decfg.set('FFT', 'dealiasing', '2/3 cython')
    np.savez_compressed(os.path.join(HERE, "transforms.npz"), **out)
    print("transforms.npz written")




def dealias_kernel_vectors():
    """Direct outputs of the reference's Cython dealias_23 (dealias_cy_2d.pyx, dealias_cy_3d.pyx), both branches: ky per row
    and ky dense over (kx, ky) / (ky, 1, kx) as the shearing box passes it."""
    import dealias_cy_2d as d2
    import dealias_cy_3d as d3
    rng = np.random.default_rng(23)
    out = {}
    # 2-D: data[x][y]
    nkx, ny = 9, 12
    kx = (np.arange(nkx, dtype=float) * 0.5).reshape(nkx, 1)
    ky0 = (np.fft.fftfreq(ny) * ny * 1.5).reshape(1, ny)
    kny = np.array([kx.max(), np.abs(ky0).max()])
    data = rng.standard_normal((nkx, ny)) + 1j * rng.standard_normal((nkx, ny))
    dense = np.ascontiguousarray(ky0 - 0.7 * kx * 1.3)
    for tag, ky in (("row", ky0), ("dense", dense)):
        o = data.copy()
        d2.dealias_23(o, kx, np.ascontiguousarray(ky), kny)
        out["d2_" + tag] = o
    out.update(d2_data=data, d2_kx=kx, d2_ky=ky0, d2_kydense=dense, d2_kny=kny)
    # 3-D: data[y][z][x]
    ny, nz, nkx = 10, 6, 5
    kx = np.arange(nkx, dtype=float).reshape(1, 1, nkx)
    ky0 = (np.fft.fftfreq(ny) * ny).reshape(ny, 1, 1)
    kz = (np.fft.fftfreq(nz) * nz * 2.0).reshape(1, nz, 1)
    kny = np.array([np.abs(ky0).max(), np.abs(kz).max(), kx.max()])
    data = rng.standard_normal((ny, nz, nkx)) + 1j * rng.standard_normal((ny, nz, nkx))
    dense = np.ascontiguousarray(ky0 - 1.1 * kx * 0.9)
    for tag, ky in (("row", ky0), ("dense", dense)):
        o = data.copy()
        d3.dealias_23(o, kx, np.ascontiguousarray(ky), kz, kny)
        out["d3_" + tag] = o
    out.update(d3_data=data, d3_kx=kx, d3_ky=ky0, d3_kydense=dense, d3_kz=kz, d3_kny=kny)
    np.savez_compressed(os.path.join(HERE, "dealias_kernels.npz"), **out)
    print("dealias_kernels.npz written")


if __name__ == "__main__":
    only = sys.argv[1:]
    for c in CASES:
        if not only or c["name"] in only:
            run_case(c)
    if not only:
        stage_kernel_vectors()
        transform_vectors()
    if not only or "dealias_kernels" in only:
        dealias_kernel_vectors()
