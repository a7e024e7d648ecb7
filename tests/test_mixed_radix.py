"""Grids that are not powers of two.  The reference transforms through FFTW / numpy.fft, which take any N, and its
own samples use such grids: 450 x 450 (samples/incompressible_hydro/2d_decaying_turbulence), 48 x 2 x 48
(samples/boussinesq_hydro/gravity_wave), 30 x 10 (swinging_wave), 100 x 100, 48^3 (samples/_deprecated).  The CUDA
library serves them with the runtime-length instantiation of the generic tile kernel (csrc/fft_core.cuh RtFac:
radix 8 / 4 / 2, hand-written 3- and 5-point butterflies, a direct DFT for any other prime factor up to 64).

CPU part: the kernel bodies through the g++ host-emulation build against the oracle (numpy.fft, the reference's own
alternative backend, representations.py:327-333); the reference-generated goldens of such grids
(tests/golden/*10x30*, *50*, *18x24*, *48x2x48*, *9x15x14*, *12x20x24*) are picked up by test_oracle_golden.py,
test_host_emulation.py and test_gpu_parity.py automatically.  GPU part: the drop-in package on the sample grids."""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "host"))
import dedalus_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def lib():
    import emul
    return emul.load()


SHAPES = [(10, 30), (30, 10), (12, 20), (100, 100), (7, 22), (45, 6), (2, 4), (6, 10, 12), (48, 2, 48), (9, 15, 14), (2, 4, 6),
          (3, 5, 26)]


@pytest.mark.parametrize("shape", SHAPES)
def test_emulated_transforms_any_length(lib, shape):
    import emul
    g = orc.Grid(shape)
    pl = emul.EmulPlan(lib, g)
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape)
    c = orc.Comp(g)
    c["xspace"] = x
    assert rel(pl.forward(x), c["kspace"]) < 2e-15
    kin = rng.standard_normal(g.kshape) + 1j * rng.standard_normal(g.kshape)
    c2 = orc.Comp(g)
    c2["kspace"] = kin.copy()
    xb, kd = pl.backward(kin)
    assert rel(xb, c2["xspace"]) < 2e-15
    assert np.array_equal(kd, c2.kdata)          # backward() masks its source in place (representations.py:347-357)


@pytest.mark.parametrize("shape", [(9, 15), (10, 6), (5, 9, 7)])
def test_emulated_transforms_without_dealiasing(lib, shape):
    """FFT.dealiasing = None zeroes the Nyquist planes only (representations.py:442-455); odd lengths have none."""
    import emul
    g = orc.Grid(shape, None, "None")
    pl = emul.EmulPlan(lib, g)
    rng = np.random.default_rng(11)
    x = rng.standard_normal(shape)
    c = orc.Comp(g)
    c["xspace"] = x
    assert rel(pl.forward(x), c["kspace"]) < 2e-15
    xb, _ = pl.backward(c["kspace"].copy())
    assert rel(xb, c["xspace"]) < 2e-15


def test_sample_grid_450(lib):
    """The 2-D decaying-turbulence sample's grid: 450 = 2 * 3^2 * 5^2."""
    import emul
    g = orc.Grid((450, 450))
    pl = emul.EmulPlan(lib, g)
    x = np.random.default_rng(450).standard_normal((450, 450))
    c = orc.Comp(g)
    c["xspace"] = x
    assert rel(pl.forward(x), c["kspace"]) < 2e-15


RHS_CASES = [("IncompressibleHydro", (30, 10), 1), ("BoussinesqHydro", (12, 20), 4), ("IncompressibleMHD", (18, 24), 2),
             ("IncompressibleHydro", (6, 10, 12), 3), ("BoussinesqHydro", (48, 2, 48), 4), ("IncompressibleMHD", (9, 15, 14), 5),
             ("IncompressibleMHD", (12, 20, 24), 5), ("IncompressibleHydro", (50, 50), 1), ("IncompressibleMHD", (10, 22, 6), 5)]


@pytest.mark.parametrize("physics,shape,cfg", RHS_CASES)
def test_emulated_rhs_any_length(lib, physics, shape, cfg):
    import emul
    kw = {"direction": "y" if len(shape) == 2 else "z"} if physics == "BoussinesqHydro" else {}
    Po = orc.PHYSICS[physics](shape, None, "2/3 cython", **kw)
    Po.parameters.update(dict(nu=1e-3, eta=1e-3, kappa=1e-3))
    do = orc.synthetic_ic(Po, cfg)
    y0 = do.kvector().copy()
    de = Po.create_fields(0.)
    Po.RHS(do, de)
    pl = emul.EmulPlan(lib, orc.Grid(shape))
    params = dict(Po.parameters)
    if kw:
        params["boussinesq_direction"] = kw["direction"]
    d, s = pl.rhs(physics, params, list(y0), flags=1 | (2 if physics == "IncompressibleMHD" else 0))
    assert rel(d, de.kvector()) < 1e-13
    assert rel(s, do.kvector()) < 1e-13


def test_unsupported_lengths_are_refused(lib):
    """A prime factor above 64 or a length above 2048 is an error with a message, not a wrong answer."""
    import ctypes as C
    lib.ddl_last_error.restype = C.c_char_p
    for n, why in ((2 * 67, b"prime factors"), (4096, b"unsupported"), (1, b"unsupported")):
        k = np.arange(n, dtype=np.float64)
        keep = np.zeros(n, dtype=np.uint8)
        keep[0] = 1
        shape = np.array([n, 16], dtype=np.int64)
        k16 = np.arange(9, dtype=np.float64)
        keep16 = (k16 < 5).astype(np.uint8)
        plan = C.c_void_p()
        rc = lib.ddl_plan_create(C.byref(plan), 2, shape.ctypes.data_as(C.c_void_p), k16.ctypes.data_as(C.c_void_p),
                                 k.ctypes.data_as(C.c_void_p), None, keep16.ctypes.data_as(C.c_void_p),
                                 keep.ctypes.data_as(C.c_void_p), None)
        assert rc != 0 and why in lib.ddl_last_error(), (n, lib.ddl_last_error())


# ------------------------------------------------------------------------------------------------ GPU
GPU_RUNS = [
    ("IncompressibleHydro", (450, 450), dict(nu=1e-3), "RK2mid", 2e-3, 2, 1),          # 2d_decaying_turbulence
    ("BoussinesqHydro", (48, 2, 48), dict(nu=1e-3, kappa=1e-3), "RK2mid", 5e-3, 3, 4),   # gravity_wave
    ("IncompressibleHydro", (30, 10), dict(nu=1e-2), "RK2mid", 1e-2, 3, 1),              # swinging_wave's grid
    ("IncompressibleHydro", (100, 100), dict(nu=1e-3), "RK4", 5e-3, 3, 1),
    ("IncompressibleMHD", (48, 48, 48), dict(nu=1e-3, eta=1e-3), "RK4", 5e-3, 2, 5),
    ("IncompressibleMHD", (9, 15, 14), dict(nu=1e-3, eta=2e-3), "RK2trap", 5e-3, 3, 5),
    ("IncompressibleMHD", (96, 60), dict(nu=1e-3, eta=1e-3), "CrankNicholsonVisc", 2e-3, 3, 2),
]


@pytest.mark.gpu
@pytest.mark.parametrize("physics,shape,params,integ,dt,nsteps,cfg", GPU_RUNS)
def test_steps_on_sample_grids_match_oracle(physics, shape, params, integ, dt, nsteps, cfg):
    from conftest import native_lib_expected
    from devutil import dev_physics, oracle_physics, set_state, get_state
    import dedalus.time_stepping.api as tapi
    import dedalus.analysis.volume_average as va
    native_lib_expected()
    Po = oracle_physics(physics, shape, None, params)
    do = orc.synthetic_ic(Po, cfg)
    P = dev_physics(physics, shape, None, params)
    data = P.create_fields(0.)
    set_state(data, do.kvector())
    to, ti = orc.INTEGRATORS[integ](Po), getattr(tapi, integ)(P)
    for _ in range(nsteps):
        to.do_advance(do, dt)
        ti.do_advance(data, dt)
    assert rel(get_state(data), do.kvector()) < 1e-10
    ref = orc.invariants(do)
    assert abs(va.ekin(data) - ref["ekin"]) < 1e-12 * max(1.0, abs(ref["ekin"]))
    assert abs(P.compute_dt(data) - Po.compute_dt(do)) < 1e-12 * Po.compute_dt(do)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(450, 450), (48, 2, 48), (7, 22), (1000, 6), (36, 50, 20)])
def test_transforms_on_any_length(shape):
    import torch
    from conftest import native_lib_expected
    from dedalus.data_objects.api import FourierRepresentation
    native_lib_expected()
    g = orc.Grid(shape)
    x = np.random.default_rng(7).standard_normal(shape)
    co = orc.Comp(g)
    co["xspace"] = x
    c = FourierRepresentation(None, shape, (2 * np.pi,) * len(shape))
    c["xspace"] = torch.from_numpy(x)
    assert rel(c["kspace"].cpu().numpy(), co["kspace"]) < 1e-14
    assert rel(c["xspace"].cpu().numpy(), co["xspace"]) < 1e-14
