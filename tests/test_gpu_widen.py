"""GPU tests of the SURVEY 8(f) rows built on device reductions: volume-average diagnostics from one
sweep (ddl_reduce_invariants), the CFL limit from inside the x pass (ddl_reduce_max_square,
ddl_rhs_capture_max) and advance(dt=None) taking its time step from the step's own first RHS.
Through the drop-in Python API -> C ABI -> kernels, against the oracle's restatement of
dedalus/analysis/volume_average.py, fields.py:153-157 and physics.py:151-158,601-610,714-721,821-836."""
import numpy as np
import pytest

from devutil import rel, dev_physics, oracle_physics, set_state, get_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _native_lib_loaded():
    from conftest import native_lib_expected
    native_lib_expected()
    yield


def close(a, b, rtol=1e-12, atol=1e-13):
    return abs(a - b) <= atol + rtol * abs(b)


def both(physics, shape, params, cfg):
    import dedalus_oracle as orc
    Po = oracle_physics(physics, shape, None, params)
    do = orc.synthetic_ic(Po, cfg)
    P = dev_physics(physics, shape, None, params)
    data = P.create_fields(0.)
    set_state(data, do.kvector())
    return Po, do, P, data


CASES = [("IncompressibleHydro", (64, 32), dict(nu=1e-3), 1), ("BoussinesqHydro", (32, 64), dict(nu=1e-3, kappa=2e-3), 4),
         ("IncompressibleMHD", (64, 64), dict(nu=1e-3, eta=1e-3), 2), ("IncompressibleHydro", (32, 32, 32), dict(nu=1e-3), 3),
         ("BoussinesqHydro", (16, 32, 64), dict(nu=1e-3, kappa=1e-3), 4), ("IncompressibleMHD", (32, 32, 32), dict(nu=1e-3, eta=2e-3), 5)]


@pytest.mark.parametrize("physics,shape,params,cfg", CASES)
def test_volume_average_tasks_match_reference_definitions(physics, shape, params, cfg):
    import dedalus_oracle as orc
    import dedalus.analysis.volume_average as va
    Po, do, P, data = both(physics, shape, params, cfg)
    ref = orc.invariants(do)
    assert close(va.ekin(data), ref["ekin"])
    assert close(va.divergence_sum(data), ref["div_sum"], atol=1e-11)
    assert abs(va.divergence(data) - ref["divergence"]) < 1e-11
    assert close(va.vort_cenk(data), ref["cenk_num"] / ref["cenk_den"])
    names = ["ux2", "uy2", "uz2"][:len(shape)]
    for j, n in enumerate(names):
        assert close(getattr(va, n)(data), ref["msq"][j])
    if len(shape) == 2:
        assert close(va.enstrophy(data), ref["enstrophy"])
    else:
        assert close(va.energy_dissipation(data), 2 * params["nu"] * ref["enstrophy"])
        assert close(va.kinetic_helicity(data), ref["hel_kin"], atol=1e-12)
    if physics == "IncompressibleMHD":
        assert close(va.emag(data), ref["e2"])
        assert close(va.mag_div_sum(data), ref["mag_div_sum"], atol=1e-11)
        assert abs(va.mag_div(data) - ref["mag_div"]) < 1e-11
        assert close(va.cross_helicity(data), ref["hel_cross"], atol=1e-12)
        assert close(va.current_squared(data), ref["current2"])
        for j, n in enumerate(["bx2", "by2", "bz2"][:len(shape)]):
            assert close(getattr(va, n)(data), ref["msq"][len(shape) + j])
        if len(shape) == 3:
            assert close(va.magnetic_helicity(data), ref["hel_mag"], atol=1e-12)
    if physics == "BoussinesqHydro":
        assert close(va.temp2(data), 2 * ref["e2"])
        assert close(va.thermal_energy_dissipation(data), params["kappa"] * ref["grad2_T"])


def test_fused_invariants_equal_the_tensor_level_route():
    """Same numbers from the one-sweep kernel and from the reference-shaped tensor operations it replaces
    (forced by hiding the standard field list), also for a state with energy outside the dealias mask."""
    import torch
    import dedalus.analysis.volume_average as va
    Po, do, P, data = both("IncompressibleMHD", (32, 32, 32), dict(nu=1e-3), 5)
    g = torch.Generator().manual_seed(3)
    for _, _, c in data.components():
        k = c["kspace"]                                 # handing the buffer out drops the 'dealiased' bit
        k += 1e-3 * torch.view_as_complex(torch.randn(k.shape + (2,), generator=g, dtype=torch.float64)).to(k.device)
    fused = [va.ekin(data), va.emag(data), va.divergence_sum(data), va.mag_div_sum(data), va.ux2(data), va.bz2(data),
             va.cross_helicity(data), va.kinetic_helicity(data)]
    saved = dict(va._PHYSICS_OF)
    va._PHYSICS_OF.clear()
    try:
        plain = [va.ekin(data), va.emag(data), va.divergence_sum(data), va.mag_div_sum(data), va.ux2(data), va.bz2(data),
                 va.cross_helicity(data), va.kinetic_helicity(data)]
    finally:
        va._PHYSICS_OF.update(saved)
    for a, b in zip(fused, plain):
        assert close(a, b, rtol=1e-12, atol=1e-11)


def test_volume_average_set_shares_one_sweep(tmp_path):
    import dedalus._lib as L
    import dedalus.analysis.volume_average as va
    Po, do, P, data = both("IncompressibleMHD", (32, 32, 32), dict(nu=1e-3), 5)
    vs = va.VolumeAverageSet(data, filename=str(tmp_path / "ts.dat"))
    for name in ("ekin", "emag", "ux2", "by2", "divergence_sum", "mag_div_sum", "energy_dissipation"):
        vs.add(name, "%10.5e")
    n0 = L.launch_count()
    vs.run()
    assert L.launch_count() - n0 == 2            # the sweep and its final reduction, for seven tasks
    line = open(str(tmp_path / "ts.dat")).read().strip().splitlines()[-1].split("\t")
    assert len(line) == 8 and abs(float(line[1]) - va.ekin(data)) < 1e-5 * abs(va.ekin(data))


@pytest.mark.parametrize("physics,shape,params,cfg", CASES)
def test_compute_dt_matches_reference_route(physics, shape, params, cfg):
    """compute_dt from the reduction inside the x pass == the reference's route (every component to x-space,
    fields.py:153-157) == the oracle; and it leaves the state where it was."""
    Po, do, P, data = both(physics, shape, params, cfg)
    y0 = get_state(data)
    dt_fused = P.compute_dt(data)
    # u, B become the spectra of their real parts (the image of the reference's x-space round trip): a
    # round-off-level change for these Hermitian-consistent states, and no transform is paid for it
    assert rel(get_state(data), y0) < 1e-15
    assert all(c._curr_space == "kspace" for _, _, c in data.components())
    P.dtlist = []
    P.set_dtlist(data)                              # outside compute_dt: field.max_square(), component by component
    dt_plain = min(P.dtlist)
    assert close(dt_fused, dt_plain, rtol=1e-13)
    assert close(dt_fused, Po.compute_dt(do), rtol=1e-13)


def test_compute_dt_dealiases_like_the_reference():
    """max_square goes through backward(), which masks the spectrum in place (representations.py:347-357)."""
    import torch
    Po, do, P, data = both("IncompressibleMHD", (32, 32, 32), dict(nu=1e-3), 5)
    for _, _, c in data.components():
        k = c["kspace"]
        k[5, 15, 3] = 0.3 + 0.1j                    # |kz index| = 15 >= 2/3 * 16: outside the mask
    do2 = Po.create_fields(0.)
    for (_, _, a), y in zip(do2.components(), get_state(data)):
        a["kspace"] = y
    dt_ref = Po.compute_dt(do2)
    assert close(P.compute_dt(data), dt_ref, rtol=1e-13)
    assert all(abs(c["kspace"][5, 15, 3].item()) == 0 for _, _, c in data.components())


@pytest.mark.parametrize("integ", ["RK2mid", "RK2trap", "RK4", "CrankNicholsonVisc"])
@pytest.mark.parametrize("physics,shape,params,cfg", [CASES[2], CASES[5], CASES[4]])
def test_advance_takes_dt_from_its_first_rhs(integ, physics, shape, params, cfg):
    """advance(data) with the CFL maxima captured in the step's first RHS == the reference's sequence
    dt = CFL * compute_dt(data); do_advance(data, dt) == the oracle doing the same."""
    import dedalus_oracle as orc
    import dedalus.time_stepping.api as tapi
    Po, do, P, data = both(physics, shape, params, cfg)
    P2 = dev_physics(physics, shape, None, params)
    data2 = P2.create_fields(0.)
    set_state(data2, do.kvector())
    ti, ti2, to = getattr(tapi, integ)(P, CFL=0.3), getattr(tapi, integ)(P2, CFL=0.3), orc.INTEGRATORS[integ](Po)
    ti.save_cadence = ti2.save_cadence = 10 ** 9
    ti.max_save_period = ti2.max_save_period = 1e300
    ti.iteration = ti2.iteration = 1                 # no snapshot at iteration 0
    ti2.fuse_cfl = False
    dt_old = np.finfo("d").max / 10.
    for step in range(3):
        dt = 0.3 * Po.compute_dt(do)
        dt = min(dt, 1.05 * dt_old)
        dt_old = dt
        to.do_advance(do, dt)
        ti.advance(data)
        ti2.advance(data2)
        assert close(ti.dt_old, dt, rtol=1e-12) and close(ti2.dt_old, dt, rtol=1e-12)
    assert rel(get_state(data), do.kvector()) < 1e-10
    assert rel(get_state(data), get_state(data2)) < 1e-11
    assert close(data.time, do.time) and ti.iteration == 4


def test_capture_costs_no_extra_transform():
    """Launch accounting of advance(data) for 3-D MHD RK4 once the fused path is on: the lazy-dt step is
    the fixed-dt step plus one stage launch (k1 evaluated unfused), while the reference's route adds the
    passes of a whole inverse pipeline."""
    import dedalus._lib as L
    import dedalus.time_stepping.api as tapi
    Po, do, P, data = both("IncompressibleMHD", (32, 32, 32), dict(nu=1e-3, eta=1e-3), 5)
    ti = tapi.RK4(P, CFL=0.3)
    ti.save_cadence, ti.max_save_period, ti.iteration = 10 ** 9, 1e300, 1
    ti.do_advance(data, 1e-3)
    ti.do_advance(data, 1e-3)
    n0 = L.launch_count(); ti.do_advance(data, 1e-3); fixed = L.launch_count() - n0
    n0 = L.launch_count(); ti.advance(data); lazy = L.launch_count() - n0
    ti.fuse_cfl = False
    n0 = L.launch_count(); ti.advance(data); plain = L.launch_count() - n0
    assert lazy == fixed + 1
    assert plain >= fixed + 3


# ---------------------------------------------------------------------------------------------
# Goldens produced by the REFERENCE's own initial-condition generators, CFL-controlled advance() loop and
# VolumeAverageSet tasks (tests/golden/make_sample_goldens.py -> tests/golden/samples/)
# ---------------------------------------------------------------------------------------------
import ast
import os

SAMPLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "samples")


def _ic_state(kind, P, data):
    import dedalus.init_cond.api as ic
    if kind == "tg3d":
        ic.taylor_green(data)
    elif kind == "turb2d":
        np.random.seed(1234)
        ic.turb_new(data, ic.mcwilliams_spec, k0=5., E0=1.)
    elif kind == "turb3d":
        np.random.seed(4321)
        ic.turb_new(data, ic.mcwilliams_spec, tot_en=0.7, k0=4., E0=1.)
    elif kind == "mit":
        ic.MIT_vortices(data)
    elif kind == "sincos":
        ic.sin_k(data["u"]["x"]["kspace"], (2, 3), ampl=0.5)
        ic.cos_k(data["u"]["y"]["kspace"], (1, 2), ampl=-1.5)
        ic.constant(data, "T", 2.5)
        ic.constant(data, "uy", -0.75)


@pytest.mark.parametrize("kind,physics,shape", [("tg3d", "IncompressibleHydro", (16, 16, 16)), ("turb2d", "IncompressibleHydro", (32, 32)),
                                                ("turb3d", "IncompressibleHydro", (16, 16, 32)), ("mit", "IncompressibleHydro", (32, 32)),
                                                ("sincos", "BoussinesqHydro", (16, 16))])
def test_initial_conditions_match_reference_generators(kind, physics, shape):
    z = np.load(os.path.join(SAMPLES, "init_cond.npz"))
    P = dev_physics(physics, shape)
    data = P.create_fields(0.)
    _ic_state(kind, P, data)
    ref = z[kind]
    got = get_state(data)
    assert np.abs(got - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()), kind


def test_old_turb_generator_and_projection_match_reference():
    """turb() + remove_compressible() (init_cond.py:343-389) against the reference's own run (seeded numpy generator)."""
    import dedalus.init_cond.api as ic
    z = np.load(os.path.join(SAMPLES, "init_cond.npz"))
    P = dev_physics("IncompressibleHydro", (24, 32))
    data = P.create_fields(0.)
    np.random.seed(99)
    ic.turb(data["u"]["x"], data["u"]["y"], ic.mcwilliams_spec, k0=4., E0=1.)
    assert np.abs(get_state(data) - z["turb_old"]).max() <= 1e-13 * np.abs(z["turb_old"]).max()
    ic.remove_compressible(data["u"]["x"], data["u"]["y"])
    assert np.abs(get_state(data) - z["turb_old_projected"]).max() <= 1e-13 * np.abs(z["turb_old_projected"]).max()


def _set_run_ic(tag, data):
    import torch
    import dedalus.init_cond.api as ic
    if tag == "cfl_turb2d":
        np.random.seed(1234)
        ic.turb_new(data, ic.mcwilliams_spec, k0=5., E0=1.)
    elif tag == "cfl_mhd3d":
        np.random.seed(77)
        ic.turb_new(data, ic.mcwilliams_spec, tot_en=0.5, k0=3., E0=1.)
        rng = np.random.default_rng(5)
        for _, c in data["B"]:
            c["xspace"] = torch.from_numpy(rng.standard_normal(tuple(int(n) for n in c.local_shape["xspace"])))
            c["kspace"]
        data["B"].div_free()
    elif tag == "cfl_bouss2d":
        ic.sin_k(data["T"]["kspace"], (1, 1), ampl=0.1)
        ic.cos_k(data["u"]["x"]["kspace"], (1, 1), ampl=0.05)
        ic.cos_k(data["u"]["y"]["kspace"], (1, 1), ampl=-0.05)
    elif tag == "cfl_bouss3d":
        rng = np.random.default_rng(8)
        for _, _, c in data.components():
            c["xspace"] = torch.from_numpy(rng.standard_normal(tuple(int(n) for n in c.local_shape["xspace"])))
            c["kspace"]
        data["u"].div_free()


@pytest.mark.parametrize("fuse_cfl", [True, False])
@pytest.mark.parametrize("tag", ["cfl_turb2d", "cfl_mhd3d", "cfl_bouss2d", "cfl_bouss3d"])
def test_cfl_controlled_runs_match_the_reference(tag, fuse_cfl):
    """The reference's own loop `while ...: ti.advance(data)` (dt from cfl_dt -> compute_dt -> max_square, 5 % growth
    cap) on its own initial conditions: same dt sequence, same final state, same VolumeAverageSet numbers."""
    import dedalus.time_stepping.api as tapi
    import dedalus.analysis.volume_average as va
    z = np.load(os.path.join(SAMPLES, tag + ".npz"))
    meta = ast.literal_eval(str(z["meta"]))
    P = dev_physics(meta["physics"], meta["shape"], None, meta["params"])
    data = P.create_fields(0.)
    _set_run_ic(tag, data)
    assert rel(get_state(data), z["y0"]) < 1e-13
    names = [str(n) for n in z["task_names"]]
    known = va.VolumeAverageSet.known_analysis
    for n, want in zip(names, z["tasks0"]):
        got = known[n](data)
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (n, got, want)
    ti = getattr(tapi, meta["integ"])(P, CFL=meta["CFL"])
    ti.save_cadence, ti.max_save_period, ti.iteration = 10 ** 9, 1e300, 1
    ti.fuse_cfl = fuse_cfl
    for want in z["dts"]:
        ti.advance(data)
        assert abs(ti.dt_old - want) <= 1e-12 * want
    assert rel(get_state(data), z["y1"]) < 1e-10
    assert abs(data.time - float(z["time"])) < 1e-13
    for n, want in zip(names, z["tasks1"]):
        got = known[n](data)
        assert abs(got - want) <= 1e-11 * max(1.0, abs(want)), (n, got, want)


# ---------------------------------------------------------------------------------------------
# States that are not solenoidal: the reference evaluates u.grad u whatever the state (physics.py:197-228);
# the package measures the compressive fraction of caller-written states and switches to the advective-form
# policies (include/ddl.h DDL_*_ADV)
# ---------------------------------------------------------------------------------------------
def _compressive_state(Po, seed):
    import dedalus_oracle as orc
    do = Po.create_fields(0.)
    rng = np.random.default_rng(seed)
    for _, _, c in do.components():
        c["xspace"] = 0.3 * rng.standard_normal(Po.g.shape)
        c.require_space("kspace")
    return do


NONSOL = [("IncompressibleHydro", (32, 64), dict(nu=1e-3)), ("BoussinesqHydro", (32, 32), dict(nu=1e-3, kappa=2e-3, g=1.5, beta=0.5)),
          ("IncompressibleMHD", (64, 32), dict(nu=1e-3, eta=2e-3, rho0=0.7)), ("IncompressibleHydro", (16, 32, 32), dict(nu=1e-2)),
          ("BoussinesqHydro", (32, 16, 32), dict(nu=1e-2, kappa=1e-2, alpha_t=0.5)), ("IncompressibleMHD", (32, 32, 16), dict(nu=1e-2, eta=1e-2)),
          # x length 128: the specialised fused x pass with the advective-form policies (hydro, MHD; Boussinesq stays generic)
          ("IncompressibleHydro", (8, 16, 128), dict(nu=1e-2)), ("IncompressibleMHD", (8, 8, 128), dict(nu=1e-2, eta=1e-2)),
          ("BoussinesqHydro", (8, 8, 128), dict(nu=1e-2, kappa=1e-2))]


@pytest.mark.parametrize("physics,shape,params", NONSOL)
def test_rhs_of_a_compressive_state_matches_the_reference_form(physics, shape, params):
    Po = oracle_physics(physics, shape, None, params)
    do = _compressive_state(Po, 17)
    P = dev_physics(physics, shape, None, params)
    data, deriv = P.create_fields(0.), P.create_fields(0.)
    set_state(data, do.kvector())
    ko = Po.create_fields(0.)
    Po.RHS(do, ko)
    P.RHS(data, deriv)
    assert not P.verify_solenoidal(data)
    assert rel(get_state(deriv), ko.kvector()) < 1e-13
    assert rel(get_state(data), do.kvector()) < 1e-14          # MHD: the state is replaced by its x-space round-trip image
    # the same state made solenoidal takes the conservative pipeline again and still matches
    for name, f in data:
        if name in ("u", "B"):
            f.div_free()
    for name, f in do:
        if f.ncomp > 1:
            f.div_free()
    Po.RHS(do, ko)
    P.RHS(data, deriv)
    assert P.verify_solenoidal(data)
    assert rel(get_state(deriv), ko.kvector()) < 1e-13


@pytest.mark.parametrize("integ", ["RK2mid", "RK2trap", "RK4", "CrankNicholsonVisc"])
@pytest.mark.parametrize("physics,shape,params", [NONSOL[2], NONSOL[3], NONSOL[4], NONSOL[5]])
def test_steps_from_a_compressive_state_match_oracle(integ, physics, shape, params):
    """The compressive part is never removed (the derivative is projected, the state is not): every stage state
    inherits it, every RHS of the run takes the advective-form policy, no stage is fused."""
    import dedalus_oracle as orc
    import dedalus.time_stepping.api as tapi
    Po = oracle_physics(physics, shape, None, params)
    do = _compressive_state(Po, 23)
    P = dev_physics(physics, shape, None, params)
    data = P.create_fields(0.)
    set_state(data, do.kvector())
    ti, to = getattr(tapi, integ)(P), orc.INTEGRATORS[integ](Po)
    for _ in range(4):
        ti.do_advance(data, 2e-3)
        to.do_advance(do, 2e-3)
    assert rel(get_state(data), do.kvector()) < 1e-10
    assert not any(c._soln for n, _, c in data.components() if n == "u")
    assert orc.divergence_sum(do, "u") > 1.0


def test_roundoff_divergence_keeps_the_conservative_pipeline():
    Po, do, P, data = both("IncompressibleMHD", (32, 32, 32), dict(nu=1e-3, eta=1e-3), 5)
    assert P.verify_solenoidal(data)
    import dedalus._lib as L
    n0 = L.launch_count()
    assert P.verify_solenoidal(data) and L.launch_count() == n0       # cached: no sweep, no sync
    data["u"]["x"]["kspace"]                                          # handing a buffer out asks for a re-check
    assert data["u"]["x"]._soln is None
    assert P.verify_solenoidal(data) and L.launch_count() == n0 + 2


# ---------------------------------------------------------------------------------------------
# Rotation, forcing and passive-tracer branches of the reference RHS (physics.py:560-586,709-711) against
# runs of the reference itself (tests/golden/samples/options.npz)
# ---------------------------------------------------------------------------------------------
OPTION_CASES = {
    "rot3d": ("IncompressibleHydro", (16, 16, 16), dict(nu=1e-2, Omega=np.array([1.0, 0.2, 0.3])), "RK2mid"),
    "rot2d": ("IncompressibleHydro", (16, 32), dict(nu=1e-2, Omega=0.7), "RK2trap"),
    "force2d": ("IncompressibleHydro", (16, 32), dict(nu=1e-2), "RK2mid"),
    "heat3d": ("BoussinesqHydro", (16, 16, 16), dict(nu=1e-2, kappa=1e-2), "RK2mid"),
    "tracer2d": ("IncompressibleHydro", (16, 32), dict(nu=1e-2, c_diff=2e-2), "RK2mid"),
    "tracer3d": ("IncompressibleHydro", (16, 16, 16), dict(nu=1e-2, c_diff=0.), "RK2trap"),
}


@pytest.mark.parametrize("tag", sorted(OPTION_CASES))
def test_rhs_options_match_reference_runs(tag):
    import torch
    import dedalus.time_stepping.api as tapi
    from dedalus.config import decfg
    z = np.load(os.path.join(SAMPLES, "options.npz"))
    physics, shape, params, integ = OPTION_CASES[tag]
    decfg.set("physics", "use_tracer", "True" if tag.startswith("tracer") else "False")
    try:
        P = dev_physics(physics, shape, None, params)
        data = P.create_fields(0.)
        if tag.startswith("tracer"):
            assert list(data.fields) == ["u", "c"]
        set_state(data, z[tag + "_y0"])
        dev = next(data.components())[2]._k.device
        if tag == "force2d":
            F = [torch.from_numpy(f).to(dev) for f in z[tag + "_F"]]
            P.set_velocity_forcing(lambda d, i: F[i])
        if tag == "heat3d":
            F = [torch.from_numpy(f).to(dev) for f in z[tag + "_F"]]
            P.set_thermal_forcing(lambda d: F[0])
        ti = getattr(tapi, integ)(P)
        for _ in range(3):
            ti.do_advance(data, 1e-2)
        assert rel(get_state(data), z[tag + "_y1"]) < 1e-10
    finally:
        decfg.set("physics", "use_tracer", "False")


TRACER_CASES = {
    "tracer_bouss3d": ("BoussinesqHydro", (16, 16, 16), dict(nu=1e-2, kappa=2e-2, c_diff=5e-3), "RK2mid"),
    "tracer_mhd3d": ("IncompressibleMHD", (16, 16, 16), dict(nu=1e-2, eta=2e-2, c_diff=1e-2), "RK2trap"),
    "tracer_mhd2d": ("IncompressibleMHD", (16, 32), dict(nu=1e-2, eta=1e-2, c_diff=0.), "RK2mid"),
}


@pytest.mark.parametrize("tag", sorted(TRACER_CASES))
def test_tracer_inherited_by_boussinesq_and_mhd_matches_reference_runs(tag):
    """In the reference the passive tracer lives in IncompressibleHydro and BoussinesqHydro / IncompressibleMHD inherit it
    (physics.py:467-470,515-522,578-583): state (u, c, T | B).  Goldens: runs of the reference itself
    (tests/golden/make_tracer_goldens.py).  Also: CFL limit, RK4 / CN and the solenoidal check work on such a state."""
    import dedalus.time_stepping.api as tapi
    from dedalus.config import decfg
    z = np.load(os.path.join(SAMPLES, "options_tracer.npz"))
    physics, shape, params, integ = TRACER_CASES[tag]
    decfg.set("physics", "use_tracer", "True")
    try:
        P = dev_physics(physics, shape, None, params)
        data = P.create_fields(0.)
        assert list(data.fields) == [str(n) for n in z[tag + "_fields"]] == ["u", "c", "T" if physics == "BoussinesqHydro" else "B"]
        set_state(data, z[tag + "_y0"])
        ti = getattr(tapi, integ)(P)
        for _ in range(3):
            ti.do_advance(data, 1e-2)
        assert rel(get_state(data), z[tag + "_y1"]) < 1e-10
        # the tracer is passive: the other fields evolve exactly as without it
        decfg.set("physics", "use_tracer", "False")
        Q = dev_physics(physics, shape, None, {k: v for k, v in params.items() if k != "c_diff"})
        plain = Q.create_fields(0.)
        keep = [j for j, (n, _, _) in enumerate(data.components()) if n != "c"]
        set_state(plain, z[tag + "_y0"][keep])
        tq = getattr(tapi, integ)(Q)
        for _ in range(3):
            tq.do_advance(plain, 1e-2)
        assert rel(get_state(plain), z[tag + "_y1"][keep]) < 1e-10
        decfg.set("physics", "use_tracer", "True")
        dt = P.compute_dt(data)
        assert abs(dt - Q.compute_dt(plain)) < 1e-12 * dt
        for other in ("RK4", "CrankNicholsonVisc"):
            t2 = getattr(tapi, other)(P)
            for _ in range(2):
                t2.do_advance(data, 2e-3)
            assert np.isfinite(get_state(data)).all()
    finally:
        decfg.set("physics", "use_tracer", "False")


@pytest.mark.parametrize("physics,shape", [("IncompressibleMHD", (16, 32, 128)), ("BoussinesqHydro", (16, 16, 256)),
                                           ("IncompressibleHydro", (16, 16, 512)), ("IncompressibleMHD", (8, 16, 512)),
                                           ("IncompressibleMHD", (8, 16, 1024))])
def test_xfused_launch_variants_agree(physics, shape):
    """ddl_set_option("xfused_variant", v): the CTA shapes 0-2, variant 3 (retained-mode count of the 2/3 rule as
    a compile-time constant) and the branch-free input packs 6 and 8 (registers only / mirrored half through shared memory)
    are the same arithmetic per pencil (bit-identical in the host emulation; the device
    build is held to round-off because the compiler may contract multiply-adds differently per instantiation)."""
    import dedalus._lib as L
    import dedalus_oracle as orc
    params = dict(nu=1e-3, eta=1e-3) if physics == "IncompressibleMHD" else dict(nu=1e-3)
    Po = oracle_physics(physics, shape, None, params)
    y0 = orc.synthetic_ic(Po, 5).kvector()
    out = []
    try:
        L.set_option("traceless_flux", 0)          # the launch variants belong to the six-product policies
        for v in (0, 1, 2, 3, 6, 8):
            L.set_option("xfused_variant", v)
            P = dev_physics(physics, shape, None, params)
            data, deriv = P.create_fields(0.), P.create_fields(0.)
            set_state(data, y0)
            P.RHS(data, deriv)
            out.append(get_state(deriv))
    finally:
        L.set_option("xfused_variant", 0)
        L.set_option("traceless_flux", 1)
    for o in out[1:]:
        assert rel(o, out[0]) < 1e-14


@pytest.mark.parametrize("physics,shape,params", [("IncompressibleHydro", (16, 24, 32), dict(nu=1e-3)), ("BoussinesqHydro", (32, 16, 16), dict(nu=1e-3, kappa=2e-3)),
                                                  ("IncompressibleMHD", (32, 32, 32), dict(nu=1e-3, eta=2e-3)), ("IncompressibleMHD", (16, 16, 128), dict(nu=1e-3, eta=2e-3)),
                                                  ("BoussinesqHydro", (16, 256, 128), dict(nu=1e-3, kappa=2e-3)), ("IncompressibleHydro", (8, 16, 512), dict(nu=1e-3))])
def test_traceless_flux_policies_agree(physics, shape, params):
    """ddl_set_option("traceless_flux", 1) (the default): the one-rank 3-D RHS forms 5 momentum products, T_ij - delta_ij T_zz,
    instead of 6 -- the part taken out is a pressure, which the solenoidal projection annihilates mode by mode
    (csrc/physics_ops.cuh Hydro3T / Bouss3T / MHD3T): one forward transform fewer through the x pass, the forward y / z passes and
    the assembly.  Same derivative as the six-product policies and as the oracle to the rounding of the projection, for the plain
    RHS, for the RHS with the CFL capture, and through three RK4 steps (fused stages)."""
    import dedalus._lib as L
    import dedalus.time_stepping.api as tapi
    import dedalus_oracle as orc
    Po = oracle_physics(physics, shape, None, params)
    do = orc.synthetic_ic(Po, 5)
    y0 = do.kvector()
    ko = Po.create_fields(0.)
    Po.RHS(do, ko)
    ref = ko.kvector()
    rhs, stepped, dts = [], [], []
    try:
        for v in (0, 1):
            L.set_option("traceless_flux", v)
            P = dev_physics(physics, shape, None, params)
            data, deriv = P.create_fields(0.), P.create_fields(0.)
            set_state(data, y0)
            P.RHS(data, deriv)
            rhs.append(get_state(deriv))
            dts.append(float(P.compute_dt(data)))
            ti = tapi.RK4(P)
            for _ in range(3):
                ti.do_advance(data, 1e-3)
            stepped.append(get_state(data))
    finally:
        L.set_option("traceless_flux", 1)
    assert np.isfinite(rhs[1]).all() and np.isfinite(stepped[1]).all()
    assert rel(rhs[1], rhs[0]) < 1e-14 and rel(rhs[1], ref) < 1e-13 and rel(rhs[0], ref) < 1e-13
    assert rel(stepped[1], stepped[0]) < 1e-14
    assert abs(dts[1] - dts[0]) <= 1e-14 * abs(dts[0])


@pytest.mark.parametrize("physics,shape", [("IncompressibleMHD", (256, 16, 16)), ("IncompressibleHydro", (16, 256, 16)), ("BoussinesqHydro", (512, 8, 16)),
                                           ("IncompressibleHydro", (8, 512, 16)), ("IncompressibleMHD", (256, 256, 16))])
def test_two_stage_strided_pass_agrees(physics, shape):
    """ddl_set_option("strided_two", 1): the y / z passes of lengths 256 and 512 as two register butterflies (16 x 16, 16 x 32)
    around ONE trip through shared memory (csrc/fast_two.cuh) instead of three radix-8 stages: same RHS as the default kernels
    to round-off and as the oracle to the usual tolerance; plain transforms (forward / backward of one component) included."""
    import torch
    import dedalus._lib as L
    import dedalus_oracle as orc
    params = dict(nu=1e-3, eta=1e-3) if physics == "IncompressibleMHD" else dict(nu=1e-3)
    Po = oracle_physics(physics, shape, None, params)
    do = orc.synthetic_ic(Po, 5)
    y0 = do.kvector()
    ko = Po.create_fields(0.)
    Po.RHS(do, ko)
    ref = ko.kvector()
    out, back = [], []
    try:
        for v in (0, 1, 2):
            L.set_option("strided_two", v)
            P = dev_physics(physics, shape, None, params)
            data, deriv = P.create_fields(0.), P.create_fields(0.)
            set_state(data, y0)
            P.RHS(data, deriv)
            out.append(get_state(deriv))
            c = data["u"]["x"]
            x = c["xspace"].clone()
            c["xspace"] = x
            back.append((x.cpu().numpy().copy(), c["kspace"].clone().cpu().numpy()))
    finally:
        L.set_option("strided_two", 1)
    for v in (1, 2):
        assert np.isfinite(out[v]).all()
        assert rel(out[v], out[0]) < 1e-13
        assert rel(back[v][0], back[0][0]) < 1e-13 and rel(back[v][1], back[0][1]) < 1e-13
        assert rel(out[v], ref) < 1e-12


@pytest.mark.parametrize("physics,shape", [("IncompressibleMHD", (16, 16, 32)), ("BoussinesqHydro", (16, 32, 16)), ("IncompressibleHydro", (16, 16, 16)),
                                           ("IncompressibleMHD", (32, 48)), ("BoussinesqHydro", (32, 32)), ("IncompressibleHydro", (48, 32))])
@pytest.mark.parametrize("stepper", ["RK4", "RK2mid", "RK2trap", "CrankNicholsonVisc"])
def test_assemble_stage_launch_variants_agree(physics, shape, stepper):
    """ddl_set_option("assemble_variant", v): the spectral assembly fused with the stage update loads its operands where the
    arithmetic needs them (0, 4) or all up front (1-3, the default 3), at four, three or two CTAs per SM -- one arithmetic:
    three steps of every integrator (the first sets the fused path up, the others run it) end in the same bits."""
    import dedalus._lib as L
    import dedalus.time_stepping.api as tapi
    import dedalus_oracle as orc
    params = dict(nu=1e-3, eta=2e-3) if physics == "IncompressibleMHD" else (dict(nu=1e-3, kappa=2e-3) if physics == "BoussinesqHydro" else dict(nu=1e-3))
    Po = oracle_physics(physics, shape, None, params)
    y0 = orc.synthetic_ic(Po, 5).kvector()
    out, launches = [], []
    try:
        for v in (0, 1, 2, 3, 4):
            L.set_option("assemble_variant", v)
            P = dev_physics(physics, shape, None, params)
            data = P.create_fields(0.)
            set_state(data, y0)
            ti = getattr(tapi, stepper)(P)
            for _ in range(3):
                ti.do_advance(data, 2e-3)
            out.append(get_state(data))
    finally:
        L.set_option("assemble_variant", 3)
    assert np.isfinite(out[0]).all()
    for o in out[1:]:
        assert rel(o, out[0]) < 1e-15


@pytest.mark.parametrize("physics,shape", [("BoussinesqHydro", (16, 16, 16)), ("IncompressibleMHD", (16, 16, 16)), ("IncompressibleHydro", (32, 32))])
def test_caller_written_states_reach_the_fused_stage_path(physics, shape):
    """After the first (unfused) step every stage of a solenoidal, dealiased run is ONE ddl_rhs_stage call -- also for states
    the caller wrote, scalar components included (a temperature field has no divergence verdict to wait for)."""
    import torch
    import dedalus._lib as L
    import dedalus.time_stepping.api as tapi
    P = dev_physics(physics, shape, None, dict(nu=1e-3))
    data = P.create_fields(0.)
    rng = np.random.default_rng(0)
    for _, f in data:
        for _, c in f:
            c["xspace"] = torch.from_numpy(rng.standard_normal(shape))
            c["kspace"]
        if f.ncomp > 1:
            f.div_free()
    ti = tapi.RK4(P)
    ti.do_advance(data, 1e-3)
    n0 = L.launch_count()
    ti.do_advance(data, 1e-3)
    per_step = L.launch_count() - n0
    n0 = L.launch_count()
    ti.fuse_stages = False
    ti.do_advance(data, 1e-3)
    unfused = L.launch_count() - n0
    assert all(c._soln is True and c._clean for c in data.comp_list())
    assert per_step < unfused and per_step <= 4 * (6 if len(shape) == 3 else 4)


@pytest.mark.parametrize("integ,nstages", [("RK2mid", 2), ("RK2trap", 2), ("RK4", 4), ("CrankNicholsonVisc", 1)])
@pytest.mark.parametrize("shape", [(32, 32), (16, 16, 16)])
@pytest.mark.parametrize("physics", ["IncompressibleHydro", "BoussinesqHydro", "IncompressibleMHD"])
def test_steady_state_launch_counts(physics, shape, integ, nstages):
    """Every physics class x dimension x integrator settles on the minimal launch sequence: per stage the passes of the
    transform pipeline (z, y, x, y, z in 3-D; y, x, y in 2-D) and ONE fused assembly + stage-update kernel; a CFL-controlled
    step (dt taken from its own first RHS evaluation) costs exactly one launch more."""
    import torch
    import dedalus._lib as L
    import dedalus.time_stepping.api as tapi
    P = dev_physics(physics, shape, None, dict(nu=1e-3))
    data = P.create_fields(0.)
    rng = np.random.default_rng(0)
    for _, f in data:
        for _, c in f:
            c["xspace"] = torch.from_numpy(rng.standard_normal(shape))
            c["kspace"]
        if f.ncomp > 1:
            f.div_free()
    ti = getattr(tapi, integ)(P, CFL=0.3)
    ti.save_cadence, ti.max_save_period, ti.iteration = 10 ** 9, 1e300, 1
    for _ in range(2):
        ti.do_advance(data, 1e-3)
    n0 = L.launch_count()
    ti.do_advance(data, 1e-3)
    fixed = L.launch_count() - n0
    n0 = L.launch_count()
    ti.advance(data)
    lazy = L.launch_count() - n0
    assert fixed == (6 if len(shape) == 3 else 4) * nstages
    assert lazy == fixed + 1


@pytest.mark.parametrize("integ", ["RK2mid", "RK2trap", "RK4", "CrankNicholsonVisc"])
@pytest.mark.parametrize("shape,junk_at", [((32, 32), (3, 20)), ((16, 16, 32), (7, 3, 2))])
def test_hydro_states_with_content_outside_the_mask_keep_the_fused_stage_kernel(integ, shape, junk_at):
    """Hydro never dealiases its state, so entries outside the 2/3 mask persist and decay by the viscous factor
    alone.  Such a state takes the fused stage kernel for its retained modes plus ddl_stage_outside for the rest: equal to the
    oracle, and the out-of-mask entries BIT-equal to the unfused full sweeps."""
    import dedalus_oracle as orc
    import dedalus._lib as L
    import dedalus.time_stepping.api as tapi
    params = dict(nu=0.05)
    Po = oracle_physics("IncompressibleHydro", shape, None, params)
    do = orc.synthetic_ic(Po, 8)
    y0 = do.kvector()
    y0[(slice(None),) + junk_at] = 0.3 - 0.1j                 # outside the mask along the first k-space axis
    assert Po.g.dealias_mask()[junk_at]
    for j, (_, _, c) in enumerate(do.components()):
        c.kdata[...] = y0[j]
    out = {}
    for fuse in (True, False):
        P = dev_physics("IncompressibleHydro", shape, None, params)
        data = P.create_fields(0.)
        set_state(data, y0)
        ti = getattr(tapi, integ)(P)
        ti.fuse_stages = fuse
        ti.do_advance(data, 5e-3)
        import dedalus.time_stepping.time_step as ts_mod
        import dedalus.physics.physics as ph_mod
        called, real = set(), L.lib

        class Spy(object):
            def __getattr__(self, name):
                called.add(name)
                return getattr(real, name)
        ts_mod.lib = ph_mod.lib = Spy()
        try:
            for _ in range(2):
                ti.do_advance(data, 5e-3)
        finally:
            ts_mod.lib = ph_mod.lib = real
        out[fuse] = (get_state(data), called, [c._clean for c in data.comp_list()])
    to = orc.INTEGRATORS[integ](Po)
    for _ in range(3):
        to.do_advance(do, 5e-3)
    assert rel(out[True][0], do.kvector()) < 1e-10 and rel(out[False][0], do.kvector()) < 1e-10
    assert {"ddl_rhs_stage", "ddl_stage_outside"} <= out[True][1] and "ddl_rhs" not in out[True][1]
    assert "ddl_rhs" in out[False][1] and not ({"ddl_rhs_stage", "ddl_stage_outside"} & out[False][1])
    assert not any(out[True][2])                               # and the state is still known to carry the entries
    idx = (slice(None),) + junk_at
    assert np.array_equal(out[True][0][idx], out[False][0][idx]) and abs(out[True][0][idx][0]) > 0.05


def test_taylor_green_as_the_reference_writes_it_takes_the_fused_path():
    """BASELINE config 1: the reference's 2-D taylor_green writes four entries into the Nyquist-kx row (init_cond.py:43-51), which
    lie outside the mask and make the full-array divergence non-zero; the solenoidal verdict looks at the retained modes only
    (nothing else enters the pipeline's products), so the run keeps the conservative pipeline and the fused stage kernel."""
    import dedalus._lib as L
    import dedalus.time_stepping.api as tapi
    from dedalus.init_cond.api import taylor_green
    P = dev_physics("IncompressibleHydro", (128, 128), None, dict(nu=0.1))
    data = P.create_fields(0.)
    taylor_green(data)
    ti = tapi.RK2mid(P)
    for _ in range(2):
        ti.do_advance(data, 1e-2)
    n0 = L.launch_count()
    ti.do_advance(data, 1e-2)
    assert L.launch_count() - n0 == 2 * (4 + 1)                # per stage: y, x, y, fused assembly + update, out-of-mask update
    assert all(c._soln for c in data.comp_list()) and not any(c._clean for c in data.comp_list())


@pytest.mark.parametrize("physics,shape", [("IncompressibleMHD", (512, 8, 16)), ("IncompressibleMHD", (8, 512, 16)),
                                           ("BoussinesqHydro", (256, 16, 32)), ("IncompressibleHydro", (16, 1024, 16)),
                                           ("IncompressibleMHD", (128, 128, 16)), ("IncompressibleHydro", (2048, 8, 16)),
                                           ("IncompressibleHydro", (8, 2048, 16))])
def test_long_y_and_z_axes_on_small_grids(physics, shape):
    """The strided pencil passes at the lengths of the headline grids (512, 1024, 2048 along y or z) on grids small enough for the
    oracle and for the host-emulation / sanitizer harness: specialised and generic kernels agree and both match the oracle."""
    from test_gpu_parity import test_generic_and_fast_kernels_agree
    test_generic_and_fast_kernels_agree(physics, shape)


@pytest.mark.parametrize("physics,shape", [("IncompressibleMHD", (16, 16, 32)), ("IncompressibleHydro", (32, 16)), ("BoussinesqHydro", (12, 20, 24))])
def test_retained_modes_only_host_transfers(physics, shape):
    """upload_retained / download_retained (ddl_copy_boxes): a dealiased spectrum crosses PCIe as its retained box only -- same
    host array as the full download bit for bit, same device array after the upload (zero elsewhere, still known-dealiased),
    same step afterwards; a spectrum with content outside the mask refuses the short download."""
    import torch
    import dedalus.time_stepping.api as tapi
    pin = torch.cuda.is_available()
    P = dev_physics(physics, shape, None, dict(nu=1e-2))
    Po = oracle_physics(physics, shape, None, dict(nu=1e-2))
    import dedalus_oracle as orc
    y0 = orc.synthetic_ic(Po, 3).kvector()
    data = P.create_fields(0.)
    set_state(data, y0)
    comps = [c for _, _, c in data.components()]
    for c in comps:
        c.dealias()
    want = [c["kspace"].cpu().clone() for c in comps]
    host = [torch.zeros(tuple(w.shape), dtype=w.dtype, pin_memory=pin) for w in want]
    for c, h, w in zip(comps, host, want):
        assert c.download_retained(h) is h
        if pin:
            torch.cuda.synchronize()
        assert torch.equal(h, w)
        full = w.numel() * w.element_size()
        assert 0 < c.retained_bytes() <= (0.36 if len(shape) == 3 else 0.5) * full
    # the device buffers get junk everywhere, then the retained upload: bit-equal to the original, zero outside, known-dealiased
    g = torch.Generator().manual_seed(1)
    for c, h, w in zip(comps, host, want):
        c["kspace"] = torch.view_as_complex(torch.randn(tuple(w.shape) + (2,), generator=g, dtype=torch.float64))
        assert not c._clean
        c.upload_retained(h)
        assert c._clean and c._curr_space == "kspace"
        assert torch.equal(c._k.cpu(), w)
    # also from x-space (stale k buffer) and then a step: same as the step from the ordinary assignment
    for c, h in zip(comps, host):
        c["xspace"]
        c.upload_retained(h)
    ti = getattr(tapi, "RK4")(P)
    ti.do_advance(data, 1e-2)
    got = get_state(data)
    P2 = dev_physics(physics, shape, None, dict(nu=1e-2))      # its own physics object: integrating factors are attached at the
    data2 = P2.create_fields(0.)                                # first RHS call of a physics object only (physics.py:535-537)
    set_state(data2, y0)
    for _, _, c in data2.components():
        c.dealias()
    getattr(tapi, "RK4")(P2).do_advance(data2, 1e-2)
    assert rel(got, get_state(data2)) < 1e-14
    # content outside the mask: the short download refuses, the wrong host layout too
    c = comps[0]
    k = c["kspace"]
    k[(0, 0, -1) if len(shape) == 3 else (-1, 0)] = 1.0        # a Nyquist-kx entry: k-space is (ky, kz, kx) in 3-D, (kx, ky) in 2-D
    with pytest.raises(ValueError):
        c.download_retained(host[0])
    with pytest.raises(ValueError):
        c.upload_retained(host[0][..., :-1])
