#!/usr/bin/env python
"""Generate tests/golden/shear/*.npz by RUNNING THE REFERENCE ITSELF (oracle/_ref, numpy-FFT backend) in a shearing box:
FourierShearRepresentation (dedalus/data_objects/representations.py:558-740) under IncompressibleHydro / BoussinesqHydro
with the reference's runnable integrators RK2mid / RK2trap.

Run in the build container only:   python tests/golden/make_shear_goldens.py
Each case stores: the x-space noise the state is built from, the initial spectra, forward / backward transforms at a
non-zero time, the sheared ky array after the steps, one RHS evaluation and the state after n steps.

ONE method of the reference is replaced before it runs, and this is why.  The reference has two back ends for this
representation.  Its production route, rev_fftw (representations.py:685-698), executes UNNORMALISED inverse FFTW plans, so
backward() returns the field itself.  Its numpy route, rev_np (:721-740), calls numpy's normalised inverses and then
multiplies `self.kdata` -- by then a dead intermediate, xdata is a separate array under the numpy method (:189-190) -- by N
"to correct numpy normalization": x-space data come out divided by N_total, every quadratic term by N_total^2, and the MHD
right-hand side, which sends the state itself through x-space (physics.py:797-815), annihilates it (|u| ~ 1e-11 after three
steps of a unit-energy field).  FFTW / MPI cannot be built here, so `rev_like_fftw` below restates rev_fftw with numpy.fft
(same operations, the normalisation FFTW's plans have) and is installed as the numpy route; everything else -- _update_k,
fwd_np, the dealiasing, the physics classes, the integrators, the Cython kernels -- is the reference's own code.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "shear")
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import build_ref  # noqa: E402

decfg, data_api, physics_api, ts = build_ref.import_ref()
from dedalus.data_objects.api import FourierShearRepresentation  # noqa: E402

import numpy.fft as npfft  # noqa: E402


def rev_like_fftw(self):
    """rev_fftw (representations.py:685-698) restated with numpy.fft: unnormalised inverse y (z) transforms with the
    transpose, the phase factor exp(-i phase_rate t), the unnormalised c2r along x."""
    shape = self.global_shape['xspace']
    if self.ndim == 2:
        k = npfft.ifft(self.kdata, axis=1) * shape[0]
    else:
        k = npfft.ifftn(self.kdata, axes=(0, 1)) * (shape[0] * shape[1])
    self._mdata[:] = np.transpose(k, [1, 0, 2][:self.ndim])
    self._mdata *= np.exp(-1j * self._phase_rate * self.sd.time)
    self.xdata[:] = npfft.irfft(self._mdata, n=int(shape[-1]), axis=-1) * shape[-1]


FourierShearRepresentation.rev_np = rev_like_fftw

CASES = [
    dict(name="hydro2d_16_rk2mid", physics="IncompressibleHydro", shape=(16, 16), length=None, S=1.5, params=dict(nu=0.01),
         integ="RK2mid", dt=2e-2, nsteps=4, t0=0.0),
    dict(name="hydro2d_10x30_rk2mid_rot", physics="IncompressibleHydro", shape=(10, 30), length=(2 * np.pi, 6 * np.pi), S=1.5,
         params=dict(nu=0.0, Omega=1.0), integ="RK2mid", dt=2e-2, nsteps=4, t0=0.0),          # swinging_wave's grid and options
    dict(name="hydro2d_32x16_rk2trap_late", physics="IncompressibleHydro", shape=(32, 16), length=None, S=-0.75,
         params=dict(nu=0.02, viscosity_order=2), integ="RK2trap", dt=1e-2, nsteps=3, t0=3.1),  # wavenumbers already wrapped
    dict(name="hydro3d_8x16x16_rk2mid", physics="IncompressibleHydro", shape=(8, 16, 16), length=None, S=1.5,
         params=dict(nu=0.01), integ="RK2mid", dt=1e-2, nsteps=3, t0=0.0),
    dict(name="bouss3d_12x8x16_rk2mid", physics="BoussinesqHydro", shape=(12, 8, 16), length=None, S=1.0,
         params=dict(nu=0.01, kappa=0.02, g=1.3, alpha_t=0.7, beta=1.1), integ="RK2mid", dt=1e-2, nsteps=3, t0=0.5),
    dict(name="mhd3d_8x16x16_rk2mid", physics="IncompressibleMHD", shape=(8, 16, 16), length=None, S=1.5,
         params=dict(nu=0.01, eta=0.02, rho0=1.3), integ="RK2mid", dt=5e-3, nsteps=3, t0=0.0),
    dict(name="mhd2d_16x32_rk2trap", physics="IncompressibleMHD", shape=(16, 32), length=None, S=-1.0,
         params=dict(nu=0.01, eta=0.0), integ="RK2trap", dt=5e-3, nsteps=3, t0=0.2),
    # samples/incompressible_hydro/swinging_wave/simulation.py: its grid, box, rotation, shear rate, initial vorticity wave,
    # integrator and time step (the first 40 of its 80 100 steps)
    dict(name="swinging_wave_sample", physics="IncompressibleHydro", shape=(30, 10), length=(2 * np.pi, 100 * 2 * np.pi), S=1.5,
         params=dict(Omega=1.), integ="RK2mid", dt=1. / 150., nsteps=40, t0=0.0, ic="vorticity_wave", mode=(0.01, 4), w_amp=0.01),
    dict(name="bouss2d_16_rk2trap", physics="BoussinesqHydro", shape=(16, 16), length=None, S=2.0,
         params=dict(nu=0.01, kappa=0.0), integ="RK2trap", dt=1e-2, nsteps=3, t0=0.0),
]


def kvec(data):
    return np.stack([c['kspace'].copy() for fn, f in data for i, c in f])


def build(c, noise):
    decfg.set('FFT', 'dealiasing', '2/3 cython')
    decfg.set('physics', 'boussinesq_direction', 'y' if len(c["shape"]) == 2 else 'z')
    RHS = getattr(physics_api, c["physics"])(c["shape"], FourierShearRepresentation, c["length"])
    RHS.parameters['shear_rate'] = c["S"]
    RHS.parameters.update(c["params"])
    data = RHS.create_fields(c["t0"])
    if c.get("ic") == "vorticity_wave":
        import dedalus.init_cond.init_cond as ic
        ic.vorticity_wave(data, c["mode"], c["w_amp"])
        return RHS, data
    j = 0
    for fn, f in data:
        for i, comp in f:
            comp['xspace'] = noise[j]
            comp['kspace']
            j += 1
        if f.ncomp > 1:
            f.div_free()
    return RHS, data


def run_case(c):
    rng = np.random.default_rng(sum(ord(ch) for ch in c["name"]))
    ncomp = len(c["shape"]) + (1 if c["physics"] == "BoussinesqHydro" else 0) + (len(c["shape"]) if c["physics"] == "IncompressibleMHD" else 0)
    noise = rng.standard_normal((ncomp,) + tuple(c["shape"]))
    out = dict(noise=noise)
    RHS, data = build(c, noise)
    out["y0"] = kvec(data)
    out["ky0"] = data['u'][0].k['y'].copy()
    # transforms at the initial time: x-space image of the state and its forward transform back
    comp = data['u'][0]
    out["u0_x"] = comp['xspace'].copy()
    out["u0_k_again"] = comp['kspace'].copy()
    # one RHS on a twin
    RHS2, d2 = build(c, noise)
    deriv = RHS2.create_fields(c["t0"])
    RHS2.RHS(d2, deriv)
    out["dy0"] = kvec(deriv)
    out["y0_after_rhs"] = kvec(d2)
    ti = getattr(ts, c["integ"])(RHS)
    for _ in range(c["nsteps"]):
        ti.do_advance(data, c["dt"])
    out["y1"] = kvec(data)
    out["ky1"] = data['u'][0].k['y'].copy()
    out["time"] = data.time
    out["dt_cfl"] = RHS.compute_dt(data)
    out["y1_after_cfl"] = kvec(data)
    # the reference's CFL-controlled loop from here: advance(data) with dt = None (time_step.py:100-107,170-179)
    ti.CFL = 0.4
    ti.save_cadence, ti.max_save_period, ti.iteration = 10 ** 9, 1e300, 1          # no snapshots (h5py is not installed)
    dts = []
    for _ in range(3):
        ti.advance(data)
        dts.append(ti.dt_old)
    out["cfl_dts"] = np.array(dts)
    out["y2"] = kvec(data)
    meta = dict(c)
    meta["length"] = list(c["length"]) if c["length"] else None
    out["meta"] = np.array(repr(meta))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, c["name"] + ".npz"), **out)
    print("%-30s |y1| = %.12e  t = %g  dt_cfl = %.6e" % (c["name"], np.linalg.norm(out["y1"]), data.time, out["dt_cfl"]))


if __name__ == "__main__":
    only = sys.argv[1:]
    for c in CASES:
        if not only or c["name"] in only:
            run_case(c)
