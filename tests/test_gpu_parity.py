"""GPU parity tests: the CUDA path (through the drop-in Python API -> C ABI -> sm_100a kernels)
against (a) golden vectors produced by running the reference itself and (b) the numpy oracle on
the same seeded inputs.  Tolerance: relative L2 <= 1e-10 on the spectral state (north star),
in practice ~1e-14."""
import glob
import os

import numpy as np
import pytest

from devutil import GOLDEN, rel, load_case, dev_physics, oracle_physics, set_state, get_state

pytestmark = pytest.mark.gpu

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
               if os.path.basename(p) not in ("stage_kernels.npz", "transforms.npz", "dealias_kernels.npz"))
DEALIASED = [c for c in CASES if "nodealias" not in c]
TOL = 1e-10


@pytest.fixture(scope="module", autouse=True)
def _native_lib_loaded():
    from conftest import native_lib_expected
    native_lib_expected()
    yield


@pytest.mark.parametrize("name,shape,L,dl", [("t2d", (16, 32), (2 * np.pi, 2 * np.pi), "2/3 cython"),
                                             ("t3d", (8, 16, 32), (2.0, 3.0, 5.0), "2/3 cython"),
                                             ("t3dn", (16, 16, 16), (2 * np.pi,) * 3, "None")])
def test_transforms_match_reference(name, shape, L, dl):
    import torch
    from dedalus.config import decfg
    from dedalus.data_objects.api import FourierRepresentation
    z = np.load(os.path.join(GOLDEN, "transforms.npz"))
    decfg.set("FFT", "dealiasing", dl)
    c = FourierRepresentation(None, shape, L)
    c["xspace"] = torch.from_numpy(z[name + "_x"])
    k = c["kspace"].cpu().numpy()
    assert rel(k, z[name + "_k"]) < 1e-14
    assert rel(c.deriv("x").cpu().numpy(), z[name + "_derivx"]) < 1e-15
    assert rel(c.deriv("y").cpu().numpy(), z[name + "_derivy"]) < 1e-15
    assert rel(c.k2().cpu().numpy(), z[name + "_k2"]) < 1e-15
    assert rel(c["xspace"].cpu().numpy(), z[name + "_xb"]) < 1e-14
    assert c.fwd_count == 1 and c.rev_count == 1
    decfg.set("FFT", "dealiasing", "2/3 cython")


@pytest.mark.parametrize("name", DEALIASED)
def test_rhs_matches_reference(name):
    z, meta = load_case(name)
    P = dev_physics(meta["physics"], meta["shape"], meta["length"], meta["params"])
    data, deriv = P.create_fields(0.), P.create_fields(0.)
    set_state(data, z["y0"])
    P.RHS(data, deriv)
    d = get_state(deriv)
    if meta["ic"] == "taylor_green":
        assert np.abs(d - z["dy0"]).max() < 1e-15      # nonlinear term is a pure gradient
    else:
        assert rel(d, z["dy0"]) < 1e-13
    assert rel(get_state(data), z["y0_after_rhs"]) < 1e-13


@pytest.mark.parametrize("name", DEALIASED)
def test_steps_match_reference(name):
    import dedalus.time_stepping.api as tapi
    import dedalus.analysis.volume_average as va
    z, meta = load_case(name)
    P = dev_physics(meta["physics"], meta["shape"], meta["length"], meta["params"])
    data = P.create_fields(0.)
    set_state(data, z["y0"])
    ti = getattr(tapi, meta["integ"])(P)
    for _ in range(meta["nsteps"]):
        ti.do_advance(data, meta["dt"])
    assert rel(get_state(data), z["y1"]) < TOL
    assert abs(data.time - float(z["time"])) < 1e-14
    inv = dict(zip([str(s) for s in z["inv_names"]], z["inv1"]))
    assert abs(va.ekin(data) - inv["ekin"]) < 1e-12
    assert abs(va.divergence_sum(data) - inv["divergence_sum"]) < 1e-11
    if "emag" in inv:
        assert abs(va.emag(data) - inv["emag"]) < 1e-12
        assert abs(va.mag_div_sum(data) - inv["mag_div_sum"]) < 1e-11


NODEALIAS = [c for c in CASES if "nodealias" in c]


@pytest.mark.parametrize("name", NODEALIAS)
def test_nodealias_runs_match_reference(name):
    """FFT.dealiasing = None or '2/3 spherical' (the goldens whose name contains "nodealias": no per-axis 2/3 rule): the
    reference's products are aliased, so the right-hand side goes through its own helper sequence (physics.Physics._unfused) instead of the fused pipeline; RHS and steps against the reference goldens."""
    import dedalus.time_stepping.api as tapi
    from dedalus.config import decfg
    z, meta = load_case(name)
    try:
        dl = meta.get("dealiasing", "None")
        P = dev_physics(meta["physics"], meta["shape"], meta["length"], meta["params"], dealiasing=dl)
        data, deriv = P.create_fields(0.), P.create_fields(0.)
        set_state(data, z["y0"])
        P.RHS(data, deriv)
        assert rel(get_state(deriv), z["dy0"]) < 1e-13
        assert rel(get_state(data), z["y0_after_rhs"]) < 1e-13
        P = dev_physics(meta["physics"], meta["shape"], meta["length"], meta["params"], dealiasing=dl)
        data = P.create_fields(0.)
        set_state(data, z["y0"])
        ti = getattr(tapi, meta["integ"])(P)
        for _ in range(meta["nsteps"]):
            ti.do_advance(data, meta["dt"])
        assert rel(get_state(data), z["y1"]) < TOL
    finally:
        decfg.set("FFT", "dealiasing", "2/3 cython")


@pytest.mark.parametrize("physics,shape,integ,dl", [("IncompressibleMHD", (16, 16, 16), "RK4", "None"),
                                                    ("IncompressibleHydro", (32, 24), "CrankNicholsonVisc", "None"),
                                                    ("BoussinesqHydro", (16, 16, 16), "RK4", "2/3 spherical")])
def test_nodealias_restated_integrators_match_oracle(physics, shape, integ, dl):
    """RK4 / CrankNicholsonVisc (restated) without the per-axis 2/3 rule: unfused RHS + the device stage kernels."""
    import dedalus_oracle as orc
    import dedalus.time_stepping.api as tapi
    from dedalus.config import decfg
    params = dict(nu=0.01, eta=0.02, kappa=0.01)
    try:
        Po = oracle_physics(physics, shape, None, params, dealiasing=dl)
        do = orc.synthetic_ic(Po, 5)
        P = dev_physics(physics, shape, None, params, dealiasing=dl)
        data = P.create_fields(0.)
        set_state(data, do.kvector())
        to, ti = orc.INTEGRATORS[integ](Po), getattr(tapi, integ)(P)
        for _ in range(3):
            to.do_advance(do, 5e-3)
            ti.do_advance(data, 5e-3)
        assert rel(get_state(data), do.kvector()) < TOL
    finally:
        decfg.set("FFT", "dealiasing", "2/3 cython")


ORACLE_RUNS = [
    # physics, shape, params, integrator, dt, nsteps, config id
    ("IncompressibleHydro", (128, 128), dict(nu=1e-3), "RK2mid", 2e-3, 10, 1),
    ("IncompressibleMHD", (64, 64), dict(nu=1e-3, eta=1e-3), "RK4", 2e-3, 10, 2),
    ("IncompressibleMHD", (128, 128), dict(), "RK4", 1e-3, 5, 2),
    ("BoussinesqHydro", (32, 64), dict(nu=1e-3, kappa=2e-3, g=1.5), "RK2trap", 2e-3, 5, 4),
    ("IncompressibleHydro", (32, 32, 32), dict(nu=1e-3), "RK4", 5e-3, 5, 3),
    ("BoussinesqHydro", (32, 32, 32), dict(nu=1e-3, kappa=1e-3), "RK4", 5e-3, 5, 4),
    ("IncompressibleMHD", (32, 32, 32), dict(nu=1e-3, eta=1e-3), "RK4", 5e-3, 5, 5),
    ("IncompressibleMHD", (32, 32, 32), dict(nu=0.1, eta=0.1), "RK4", 2e-2, 5, 5),      # stiff: exp branch
    ("IncompressibleMHD", (32, 32, 32), dict(), "RK4", 5e-3, 5, 5),                      # inviscid: Euler branch
    ("IncompressibleMHD", (64, 64, 64), dict(nu=1e-3, eta=1e-3), "RK2mid", 2e-3, 3, 5),
    ("IncompressibleMHD", (16, 64, 32), dict(nu=1e-3, eta=2e-3, rho0=0.5), "CrankNicholsonVisc", 2e-3, 5, 5),
    ("IncompressibleHydro", (64, 64), dict(nu=0.05), "CrankNicholsonVisc", 5e-3, 5, 1),
]


@pytest.mark.parametrize("physics,shape,params,integ,dt,nsteps,cfg", ORACLE_RUNS)
def test_steps_match_oracle(physics, shape, params, integ, dt, nsteps, cfg):
    import dedalus_oracle as orc
    import dedalus.time_stepping.api as tapi
    Po = oracle_physics(physics, shape, None, params)
    do = orc.synthetic_ic(Po, cfg)
    y0 = do.kvector()
    P = dev_physics(physics, shape, None, params)
    data = P.create_fields(0.)
    set_state(data, y0)
    to, ti = orc.INTEGRATORS[integ](Po), getattr(tapi, integ)(P)
    for _ in range(nsteps):
        to.do_advance(do, dt)
        ti.do_advance(data, dt)
    assert rel(get_state(data), do.kvector()) < TOL
    assert abs(data.time - do.time) < 1e-13


def test_taylor_green_exact_decay():
    """2-D Taylor-Green: every mode decays as exp(-2 nu t) (init_cond.py:33-51)."""
    import dedalus.time_stepping.api as tapi
    from dedalus.init_cond.api import taylor_green
    nu, dt, n = 0.1, 1e-2, 50
    P = dev_physics("IncompressibleHydro", (128, 128), None, dict(nu=nu))
    data = P.create_fields(0.)
    taylor_green(data)
    ti = tapi.RK2mid(P)
    for _ in range(n):
        ti.do_advance(data, dt)
    got = data["u"]["x"]["kspace"][1, 1].item()
    assert abs(got - (-1j / 4.0) * np.exp(-2 * nu * dt * n)) < 1e-14


def test_rk4_properties_256cubed_roundtrip():
    """Size-independent properties at a size the oracle cannot reach quickly: forward(backward(k))
    is the identity on the dealiased subspace, div u stays at round-off, B stays solenoidal."""
    import torch
    import dedalus.time_stepping.api as tapi
    import dedalus.analysis.volume_average as va
    P = dev_physics("IncompressibleMHD", (128, 128, 128), None, dict(nu=1e-3, eta=1e-3))
    data = P.create_fields(0.)
    g = torch.Generator(device="cuda").manual_seed(5)
    for _, f in data:
        for _, c in f:
            c["xspace"] = torch.randn(128, 128, 128, dtype=torch.float64, device="cuda", generator=g)
            c["kspace"]
        f.div_free()
    c = data["u"]["x"]
    k0 = c["kspace"].clone()
    c["xspace"]; k1 = c["kspace"]
    assert (k1 - k0).norm() / k0.norm() < 1e-14
    e0 = va.ekin(data) + va.emag(data)
    ti = tapi.RK4(P)
    for _ in range(3):
        ti.do_advance(data, 1e-4)
    assert va.divergence_sum(data) / 128 ** 3 < 1e-12
    assert va.mag_div_sum(data) / 128 ** 3 < 1e-12
    e1 = va.ekin(data) + va.emag(data)
    assert abs(e1 - e0) / e0 < 1e-2 and e1 < e0


@pytest.mark.parametrize("physics,shape", [("IncompressibleMHD", (32, 64, 32)),
                                           ("IncompressibleMHD", (16, 32, 128)), ("BoussinesqHydro", (32, 16, 256)),
                                           ("IncompressibleHydro", (16, 16, 512)), ("IncompressibleMHD", (16, 16, 512)),
                                           ("IncompressibleMHD", (8, 16, 1024))])
def test_generic_and_fast_kernels_agree(physics, shape):
    """The specialised sm_100a kernels (strided_fast, xfused_kernel for nx in 128..1024) and the
    generic tile kernel implement the same maths, and both match the oracle."""
    import dedalus._lib as L
    import dedalus_oracle as orc
    params = dict(nu=1e-3, eta=1e-3) if physics == "IncompressibleMHD" else dict(nu=1e-3)
    Po = oracle_physics(physics, shape, None, params)
    y0 = orc.synthetic_ic(Po, 5).kvector()
    out = []
    for fast in (1, 0):
        L.set_option("fast_kernels", fast)
        P = dev_physics(physics, shape, None, params)
        data, deriv = P.create_fields(0.), P.create_fields(0.)
        set_state(data, y0)
        P.RHS(data, deriv)
        out.append(get_state(deriv))
    L.set_option("fast_kernels", 1)
    assert rel(out[0], out[1]) < 1e-14
    do, ko = Po.create_fields(0.), Po.create_fields(0.)
    for j, (_, _, c) in enumerate(do.components()):
        c["kspace"] = y0[j]
    Po.RHS(do, ko)
    assert rel(out[0], ko.kvector()) < 1e-13


def test_config2_orszag_tang_512_rk4_vs_oracle():
    """BASELINE config 2 at full size: 2-D MHD Orszag-Tang vortex 512^2, RK4, 2/3 dealiasing."""
    import torch
    import dedalus_oracle as orc
    import dedalus.time_stepping.api as tapi
    import dedalus.analysis.volume_average as va
    params = dict(nu=1e-3, eta=1e-3)
    Po = oracle_physics("IncompressibleMHD", (512, 512), None, params)
    do = orc.orszag_tang(Po.create_fields(0.))
    P = dev_physics("IncompressibleMHD", (512, 512), None, params)
    data = P.create_fields(0.)
    set_state(data, do.kvector())
    to, ti = orc.RK4(Po), tapi.RK4(P)
    for _ in range(5):
        to.do_advance(do, 2e-3)
        ti.do_advance(data, 2e-3)
    assert rel(get_state(data), do.kvector()) < TOL
    assert abs(va.ekin(data) - orc.energy(do, "u")) < 1e-12 and abs(va.emag(data) - orc.energy(do, "B")) < 1e-12


def test_mhd_128cubed_rk4_step_vs_oracle():
    """One RK4 step of 3-D MHD at 128^3 (the size of the bounded CPU baseline) against the oracle."""
    import dedalus_oracle as orc
    import dedalus.time_stepping.api as tapi
    params = dict(nu=1e-3, eta=1e-3)
    Po = oracle_physics("IncompressibleMHD", (128, 128, 128), None, params)
    do = orc.synthetic_ic(Po, 5)
    P = dev_physics("IncompressibleMHD", (128, 128, 128), None, params)
    data = P.create_fields(0.)
    set_state(data, do.kvector())
    orc.RK4(Po).do_advance(do, 2e-3)
    tapi.RK4(P).do_advance(data, 2e-3)
    assert rel(get_state(data), do.kvector()) < TOL


@pytest.mark.parametrize("physics,n,params", [("IncompressibleHydro", 256, dict(nu=1e-3)),
                                              ("BoussinesqHydro", 512, dict(nu=1e-3, kappa=1e-3)),
                                              ("IncompressibleMHD", 512, dict(nu=1e-3, eta=1e-3))])
def test_full_size_properties(physics, n, params):
    """BASELINE configs 3-5 at full single-GPU size, through properties that need no oracle:
    transform round trip, linearity of the transform, Hermitian symmetry of the kx = 0 plane,
    solenoidal u (and B) and decaying energy after RK4 steps, zero outside the dealias mask."""
    import torch
    import dedalus.time_stepping.api as tapi
    import dedalus.analysis.volume_average as va
    P = dev_physics(physics, (n, n, n), None, params)
    data = P.create_fields(0.)
    g = torch.Generator(device="cuda").manual_seed(7)
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
    c = data["u"][0]
    k0 = c["kspace"].clone()
    x0 = c["xspace"].clone()
    k1 = c["kspace"]
    assert float((k1 - k0).norm() / k0.norm()) < 1e-13                      # round trip
    c["xspace"] = 2.5 * x0
    assert float((c["kspace"] - 2.5 * k0).norm() / k0.norm()) < 1e-13       # linearity
    c["kspace"] = k0
    del x0, k1
    c._xdata = None
    torch.cuda.empty_cache()
    e0 = va.ekin(data)
    ti = tapi.RK4(P)
    umax = float(data["u"].max_square()) ** 0.5
    for _, cc in data["u"]:
        cc["kspace"]
        cc._xdata = None
    dt = 0.2 * (2 * np.pi / n) / umax
    for _ in range(2):
        ti.do_advance(data, dt)
    assert va.divergence_sum(data) / n ** 3 < 1e-12
    if physics == "IncompressibleMHD":
        assert va.mag_div_sum(data) / n ** 3 < 1e-12
    e1 = va.ekin(data)
    assert np.isfinite(e1) and (physics != "IncompressibleHydro" or e1 < e0)
    k = data["u"][0]["kspace"]
    keep = None
    for name, kv in data["u"][0].k.items():
        kn = float(data["u"][0].kny[data["u"][0].ktrans[name]])
        m = kv.abs() < 2.0 / 3.0 * kn
        keep = m if keep is None else (keep & m)
    assert float(k[~keep.expand_as(k)].abs().max()) == 0.0                  # dealiased
    plane = k[:, :, 0]
    mirror = plane[1:, 1:].flip(0, 1).conj()
    assert float((plane[1:, 1:] - mirror).abs().max()) < 1e-12 * float(plane.abs().max())   # Hermitian kx = 0 plane


@pytest.mark.parametrize("physics,shape,params", [("IncompressibleMHD", (32, 32, 64), dict(nu=1e-3, eta=0.2)),
                                                  ("BoussinesqHydro", (32, 32, 32), dict(nu=1e-3, kappa=1e-3)),
                                                  ("IncompressibleMHD", (64, 64), dict(nu=1e-3, eta=1e-3))])
def test_rk4_fused_assembly_equals_unfused(physics, shape, params):
    """RK4 with the spectral assembly fused into the stage update (ddl_rhs_stage, taken from the second
    step on when everything is dealiased) == RHS + ddl_rk4_stage, and both match the oracle."""
    import dedalus_oracle as orc
    import dedalus.time_stepping.api as tapi
    Po = oracle_physics(physics, shape, None, params)
    do = orc.synthetic_ic(Po, 9)
    y0 = do.kvector()
    res = []
    for fuse in (True, False):
        P = dev_physics(physics, shape, None, params)
        data = P.create_fields(0.)
        set_state(data, y0)
        ti = tapi.RK4(P)
        ti.fuse_stages = fuse
        for _ in range(4):
            ti.do_advance(data, 3e-3)
        res.append(get_state(data))
        assert abs(data.time - 4 * 3e-3) < 1e-14
    assert rel(res[0], res[1]) < 1e-14
    to = orc.RK4(Po)
    for _ in range(4):
        to.do_advance(do, 3e-3)
    assert rel(res[0], do.kvector()) < TOL


@pytest.mark.parametrize("physics,shape,integ", [("IncompressibleMHD", (512, 512), "RK4"), ("IncompressibleHydro", (128, 128), "RK2mid"),
                                                 ("IncompressibleMHD", (32, 32, 32), "RK4")])
def test_cuda_graph_replay_equals_eager_steps(physics, shape, integ):
    """do_advance_graph (one captured step replayed) == the same number of eager do_advance calls."""
    import dedalus_oracle as orc
    import dedalus.time_stepping.api as tapi
    params = dict(nu=1e-3, eta=1e-3) if physics == "IncompressibleMHD" else dict(nu=1e-3)
    Po = oracle_physics(physics, shape, None, params)
    y0 = orc.synthetic_ic(Po, 2).kvector()
    res = []
    for graph in (True, False):
        P = dev_physics(physics, shape, None, params)
        data = P.create_fields(0.)
        set_state(data, y0)
        ti = getattr(tapi, integ)(P)
        for _ in range(6):
            (ti.do_advance_graph if graph else ti.do_advance)(data, 1e-3)
        assert ti.iteration == 6 and abs(data.time - 6e-3) < 1e-14
        res.append(get_state(data))
    assert rel(res[0], res[1]) < 1e-14


def test_dealiased_flag_is_reestablished_and_junk_is_kept():
    """States handed to the caller lose their 'dealiased' bit; the integrators re-check it on the device.
    A dealiased hydro state then takes the retained-only / fused path; a state with content outside
    the mask (which hydro never removes) keeps evolving it by the viscous factor, as the
    reference does.  Both must match the oracle."""
    import torch
    import dedalus_oracle as orc
    import dedalus.time_stepping.api as tapi
    import dedalus._lib as L
    shape, params = (32, 32, 32), dict(nu=0.05)
    for junk in (False, True):
        Po = oracle_physics("IncompressibleHydro", shape, None, params)
        do = orc.synthetic_ic(Po, 8)
        y0 = do.kvector()
        if junk:
            y0[:, 14, 3, 2] = 0.3 - 0.1j          # |ky| = 14 >= 2/3 * 16: outside the mask
            for j, (_, _, c) in enumerate(do.components()):
                c.kdata[...] = y0[j]
        P = dev_physics("IncompressibleHydro", shape, None, params)
        data = P.create_fields(0.)
        set_state(data, y0)
        comps = [c for _, _, c in data.components()]
        for c in comps:
            c.kdata                                  # hand the buffers out: knowledge dropped
        assert not any(c._clean for c in comps)
        ti, to = tapi.RK4(P), orc.RK4(Po)
        n0 = L.launch_count()
        for _ in range(3):
            ti.do_advance(data, 5e-3)
            to.do_advance(do, 5e-3)
        assert all(c._clean for c in comps) == (not junk)
        assert rel(get_state(data), do.kvector()) < TOL
        if junk:
            assert abs(get_state(data)[0, 14, 3, 2]) > 0.05     # still there, only damped


@pytest.mark.parametrize("integ", ["RK2mid", "RK2trap", "CrankNicholsonVisc"])
@pytest.mark.parametrize("physics,shape,params", [("IncompressibleMHD", (32, 32, 32), dict(nu=1e-3, eta=0.2)),
                                                  ("BoussinesqHydro", (32, 64), dict(nu=1e-3, kappa=2e-3))])
def test_fused_stage_integrators_equal_unfused_and_oracle(integ, physics, shape, params):
    """RK2mid / RK2trap / CN with the assembly fused into their stage updates (from the second step on)
    == the unfused launches, and both match the oracle (RK2mid / RK2trap are pinned by the reference)."""
    import dedalus_oracle as orc
    import dedalus.time_stepping.api as tapi
    Po = oracle_physics(physics, shape, None, params)
    do = orc.synthetic_ic(Po, 12)
    y0 = do.kvector()
    res = []
    for fuse in (True, False):
        P = dev_physics(physics, shape, None, params)
        data = P.create_fields(0.)
        set_state(data, y0)
        ti = getattr(tapi, integ)(P)
        ti.fuse_stages = fuse
        for _ in range(4):
            ti.do_advance(data, 2e-3)
        res.append(get_state(data))
    assert rel(res[0], res[1]) < 1e-14
    to = orc.INTEGRATORS[integ](Po)
    for _ in range(4):
        to.do_advance(do, 2e-3)
    assert rel(res[0], do.kvector()) < TOL


def test_unfused_helpers_reproduce_the_fused_mhd_rhs():
    """The reference's formulation of the 3-D MHD right-hand side (physics.py:770-819), written with the
    unfused helpers kept for analysis scripts (XgradY, curlX, XcrossY, pressure_projection: one transform
    API call per field, 36 transforms), equals the fused pipeline (15 transform-equivalents, 6 launches)."""
    import dedalus_oracle as orc
    shape, params = (32, 16, 32), dict(nu=1e-3, eta=1e-3, rho0=0.7)
    Po = oracle_physics("IncompressibleMHD", shape, None, params)
    y0 = orc.synthetic_ic(Po, 4).kvector()
    P = dev_physics("IncompressibleMHD", shape, None, params)
    data, fused, deriv = P.create_fields(0.), P.create_fields(0.), P.create_fields(0.)
    set_state(data, y0)
    P.RHS(data, fused)
    aux = P.aux_fields
    u, B = data["u"], data["B"]
    P.XgradY(u, u, aux["mathscalar"], aux["mathvector"], deriv["u"])
    for i in range(3):
        deriv["u"][i]["kspace"].mul_(-1.0)
    P.curlX(B, aux["mathvector"])
    P.XcrossY(aux["mathvector"], B, aux["mathvector2"])
    fpr = 4 * np.pi * params["rho0"]
    for i in range(3):
        deriv["u"][i]["kspace"].add_(aux["mathvector2"][i]["kspace"] / fpr)
    P.pressure_projection(data, deriv)
    P.XcrossY(u, B, aux["mathvector"])
    P.curlX(aux["mathvector"], deriv["B"])
    assert rel(get_state(deriv), get_state(fused)) < 1e-12
    assert rel(get_state(data), y0) < 1e-13          # the helpers leave the (dealiased) state intact


@pytest.mark.parametrize("shape", [(16, 512, 512), (512, 16, 512)])
def test_opt_in_async_staging_kernels_reproduce_the_default_ones(shape):
    """The two async-copy variants built in round 2 and kept opt-in because they measured slower: the persistent
    x pass with cp.async.bulk + mbarrier staging (xfused_variant = 4) and the persistent strided pass with cp.async staging
    (strided_staged = 1).  Same butterflies in the same order: the staged strided pass is bit-identical with the three-stage
    strided pass it was derived from (strided_two = 0; the default since then is the two-stage pass of csrc/fast_two.cuh, a
    different factorisation: round-off), the x pass to round-off."""
    import dedalus._lib as L
    params = dict(nu=1e-3, eta=1e-3)
    P = dev_physics("IncompressibleMHD", shape, None, params)
    data, deriv = P.create_fields(0.), P.create_fields(0.)
    import torch
    g = torch.Generator(device="cpu").manual_seed(3)
    for _, f in data:
        for _, c in f:
            c["xspace"] = torch.randn(*shape, dtype=torch.float64, generator=g)
            c["kspace"]
        f.div_free()
    out = {}
    defaults = {"strided_two": 1, "traceless_flux": 1}
    six = {"traceless_flux": 0}        # the x-pass launch variants exist for the six-product policies (csrc/tile_inst.cu)
    for tag, opts in (("default", {}), ("three_stage", {"strided_two": 0}), ("staged", {"strided_staged": 1}),
                      ("persist", dict(six, xfused_variant=4)), ("rot", dict(six, xfused_variant=5))):
        for k, v in opts.items():
            L.set_option(k, v)
        try:
            P.RHS(data, deriv)
            out[tag] = get_state(deriv)
        finally:
            for k in opts:
                L.set_option(k, defaults.get(k, 0))
    assert np.array_equal(out["staged"], out["three_stage"])
    assert rel(out["three_stage"], out["default"]) < 1e-14
    assert rel(out["persist"], out["default"]) < 1e-14 and rel(out["rot"], out["default"]) < 1e-14
    assert np.isfinite(out["default"]).all() and np.abs(out["default"]).max() > 0

