"""CPU: pin the numpy oracle (oracle/dedalus_oracle.py) against vectors produced by running the
reference itself (tests/golden/make_golden.py -> oracle/_ref)."""
import ast
import glob
import os

import numpy as np
import pytest

import dedalus_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
               if os.path.basename(p) not in ("stage_kernels.npz", "transforms.npz", "dealias_kernels.npz"))


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = ast.literal_eval(str(z["meta"]))
    return z, meta


def make_physics(meta):
    kw = {}
    if meta["physics"] == "BoussinesqHydro":
        kw["direction"] = "y" if len(meta["shape"]) == 2 else "z"
    P = orc.PHYSICS[meta["physics"]](meta["shape"], meta["length"], meta.get("dealiasing", "2/3 cython"), **kw)
    P.parameters.update(meta["params"])
    return P


def set_state(data, y):
    for j, (_, _, c) in enumerate(data.components()):
        c["kspace"] = y[j]


@pytest.mark.parametrize("name", CASES)
def test_oracle_ic_matches_reference(name):
    z, meta = load_case(name)
    P = make_physics(meta)
    data = P.create_fields(0.0)
    if meta["ic"] == "synthetic":
        orc.synthetic_ic(P, meta["cfg"], data)
    elif meta["ic"] == "taylor_green":
        orc.taylor_green(data)
    else:
        orc.orszag_tang(data)
    assert rel(data.kvector(), z["y0"]) < 1e-14


@pytest.mark.parametrize("name", CASES)
def test_oracle_rhs_matches_reference(name):
    z, meta = load_case(name)
    P = make_physics(meta)
    data, deriv = P.create_fields(0.0), P.create_fields(0.0)
    set_state(data, z["y0"])
    P.RHS(data, deriv)
    assert rel(deriv.kvector(), z["dy0"]) < 1e-13
    # F5/F7: MHD round-trips (dealiases) the state in place, hydro leaves it untouched
    assert rel(data.kvector(), z["y0_after_rhs"]) < 1e-13


@pytest.mark.parametrize("name", CASES)
def test_oracle_steps_match_reference(name):
    z, meta = load_case(name)
    P = make_physics(meta)
    data = P.create_fields(0.0)
    set_state(data, z["y0"])
    ti = orc.INTEGRATORS[meta["integ"]](P)
    for _ in range(meta["nsteps"]):
        ti.do_advance(data, meta["dt"])
    assert rel(data.kvector(), z["y1"]) < 1e-12
    assert abs(data.time - float(z["time"])) < 1e-14
    names = [str(s) for s in z["inv_names"]]
    inv = dict(zip(names, z["inv1"]))
    assert abs(orc.energy(data, "u") - inv["ekin"]) < 1e-13
    assert abs(orc.divergence_sum(data, "u") - inv["divergence_sum"]) < 1e-12
    if "emag" in inv:
        assert abs(orc.energy(data, "B") - inv["emag"]) < 1e-13
        assert abs(orc.divergence_sum(data, "B") - inv["mag_div_sum"]) < 1e-12


@pytest.mark.parametrize("nd", [2, 3])
def test_oracle_stage_kernels_match_reference_cython(nd):
    z = np.load(os.path.join(GOLDEN, "stage_kernels.npz"))
    p = "k%dd_" % nd
    s, d1, d2, IF, dt = z[p + "start"], z[p + "d1"], z[p + "d2"], z[p + "IF"], float(z[p + "dt"])
    Z = IF * dt
    assert (Z == 0).any() and ((np.abs(Z) < 0.5) & (Z != 0)).any() and (np.abs(Z) >= 0.5).any()
    o = np.zeros_like(s)
    orc.euler(s, o, d1, dt); assert rel(o, z[p + "euler"]) < 1e-15
    orc.etd1(s, o, d1, IF, dt); assert rel(o, z[p + "etd1"]) < 1e-15
    orc.etd2rk1(s, o, d1, d2, IF, dt); assert rel(o, z[p + "etd2rk1"]) < 1e-15
    orc.etd2rk2(s, o, d1, d2, IF, dt); assert rel(o, z[p + "etd2rk2"]) < 1e-15


@pytest.mark.parametrize("name,shape,L,dl", [("t2d", (16, 32), (2 * np.pi, 2 * np.pi), "2/3 cython"),
                                             ("t3d", (8, 16, 32), (2.0, 3.0, 5.0), "2/3 cython"),
                                             ("t3dn", (16, 16, 16), (2 * np.pi,) * 3, "None")])
def test_oracle_transforms_match_reference(name, shape, L, dl):
    z = np.load(os.path.join(GOLDEN, "transforms.npz"))
    g = orc.Grid(shape, L, dl)
    c = orc.Comp(g)
    c["xspace"] = z[name + "_x"]
    assert rel(c["kspace"], z[name + "_k"]) < 1e-15
    assert rel(c.deriv("x"), z[name + "_derivx"]) < 1e-15
    assert rel(c.deriv("y"), z[name + "_derivy"]) < 1e-15
    assert rel(g.k2(), z[name + "_k2"]) < 1e-15
    assert rel(c["xspace"], z[name + "_xb"]) < 1e-14


# ---- goldens of the reference's own CFL-controlled advance() loop, IC generators and volume-average tasks
# (tests/golden/make_sample_goldens.py -> tests/golden/samples/)
SAMPLES = os.path.join(GOLDEN, "samples")
TASK_KEY = {"ekin": "ekin", "emag": "e2", "enstrophy": "enstrophy", "divergence_sum": "div_sum", "mag_div_sum": "mag_div_sum",
            "divergence": "divergence", "mag_div": "mag_div"}


def oracle_task(P, data, name):
    inv = orc.invariants(data)
    nd = P.g.ndim
    if name in TASK_KEY:
        return inv[TASK_KEY[name]]
    if name in ("ux2", "uy2", "uz2"):
        return inv["msq"]["xyz".index(name[1])]
    if name in ("bx2", "by2", "bz2"):
        return inv["msq"][nd + "xyz".index(name[1])]
    if name == "temp2":
        return 2 * inv["e2"]
    if name == "vort_cenk":
        return inv["cenk_num"] / inv["cenk_den"]
    if name == "energy_dissipation":
        return 2 * P.parameters["nu"] * inv["enstrophy"]
    if name == "thermal_energy_dissipation":
        return P.parameters["kappa"] * inv["grad2_T"]
    raise KeyError(name)


@pytest.mark.parametrize("tag", ["cfl_turb2d", "cfl_mhd3d", "cfl_bouss2d", "cfl_bouss3d"])
def test_oracle_cfl_runs_and_tasks_match_reference(tag):
    z = np.load(os.path.join(SAMPLES, tag + ".npz"))
    meta = ast.literal_eval(str(z["meta"]))
    meta["length"] = None
    P = make_physics(meta)
    data = P.create_fields(0.0)
    set_state(data, z["y0"])
    names = [str(n) for n in z["task_names"]]
    for n, want in zip(names, z["tasks0"]):
        assert abs(oracle_task(P, data, n) - want) <= 1e-12 * max(1.0, abs(want)), n
    ti = orc.INTEGRATORS[meta["integ"]](P)
    dt_old = np.finfo("d").max / 10.0
    for want in z["dts"]:
        dt = min(meta["CFL"] * P.compute_dt(data), 1.05 * dt_old)      # time_step.py:170-179
        dt_old = dt
        assert abs(dt - want) <= 1e-12 * want
        ti.do_advance(data, dt)
    assert rel(data.kvector(), z["y1"]) < 1e-12
    for n, want in zip(names, z["tasks1"]):
        assert abs(oracle_task(P, data, n) - want) <= 1e-12 * max(1.0, abs(want)), n


def test_oracle_taylor_green_3d_matches_reference():
    z = np.load(os.path.join(SAMPLES, "init_cond.npz"))
    P = orc.IncompressibleHydro((16, 16, 16))
    data = P.create_fields(0.0)
    orc.taylor_green(data)
    assert np.array_equal(data.kvector(), z["tg3d"])
