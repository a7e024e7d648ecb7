"""CPU: run the CUDA library's kernel bodies through the g++ host-emulation build
(tests/host/emul.py) and compare with the reference goldens / the oracle.  This checks the
index logic of every pass (pruned tables, Hermitian pair packing, scrambled FFT order,
spectral assembly, stage updates) without a GPU; the -m gpu tests repeat it on the device."""
import ast
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "host"))
import dedalus_oracle as orc
import emul

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def lib():
    return emul.load()


def test_fft_core_host_emulation(tmp_path):
    exe = str(tmp_path / "emul_fft")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "dedalus-1.0_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host", "emul_fft.cpp"), "-o", exe])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stdout


@pytest.mark.parametrize("name,shape,L,dl", [("t2d", (16, 32), (2 * np.pi, 2 * np.pi), "2/3 cython"),
                                             ("t3d", (8, 16, 32), (2.0, 3.0, 5.0), "2/3 cython"),
                                             ("t3dn", (16, 16, 16), (2 * np.pi,) * 3, "None")])
def test_emulated_transforms_match_reference(lib, name, shape, L, dl):
    z = np.load(os.path.join(GOLDEN, "transforms.npz"))
    g = orc.Grid(shape, L, dl)
    pl = emul.EmulPlan(lib, g)
    k = pl.forward(z[name + "_x"])
    assert rel(k, z[name + "_k"]) < 1e-14
    x, kd = pl.backward(z[name + "_k"])
    assert rel(x, z[name + "_xb"]) < 1e-14
    assert rel(pl.deriv(z[name + "_k"], 0), z[name + "_derivx"]) < 1e-15
    assert rel(pl.deriv(z[name + "_k"], 1), z[name + "_derivy"]) < 1e-15


def test_emulated_backward_dealiases_source_in_place(lib):
    g = orc.Grid((16, 16, 16))
    pl = emul.EmulPlan(lib, g)
    rng = np.random.default_rng(3)
    k = rng.standard_normal(g.kshape) + 1j * rng.standard_normal(g.kshape)
    c = orc.Comp(g)
    c["kspace"] = k
    xr = c["xspace"].copy()           # numpy irfftn semantics for non-Hermitian junk too
    x, kd = pl.backward(k)
    assert np.array_equal(kd, c.kdata)
    assert rel(x, xr) < 1e-14


CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
               if os.path.basename(p) not in ("stage_kernels.npz", "transforms.npz", "dealias_kernels.npz"))


@pytest.mark.parametrize("name", [c for c in CASES if "nodealias" not in c])
def test_emulated_rhs_matches_reference(lib, name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = ast.literal_eval(str(z["meta"]))
    g = orc.Grid(meta["shape"], meta["length"], meta.get("dealiasing", "2/3 cython"))
    pl = emul.EmulPlan(lib, g)
    params = dict(meta["params"])
    if meta["physics"] == "BoussinesqHydro":
        params["boussinesq_direction"] = "y" if g.ndim == 2 else "z"
    mhd = meta["physics"] == "IncompressibleMHD"
    d, s = pl.rhs(meta["physics"], params, list(z["y0"]), flags=1 | (2 if mhd else 0))
    if meta["ic"] == "taylor_green":
        # nonlinear term is a pure gradient: compare absolutely
        assert np.abs(d - z["dy0"]).max() < 1e-15
    else:
        assert rel(d, z["dy0"]) < 1e-13
    assert rel(s, z["y0_after_rhs"]) < 1e-13


@pytest.mark.parametrize("shape", [(16, 32), (8, 16, 16)])
def test_emulated_stage_kernels_match_oracle(lib, shape):
    g = orc.Grid(shape, None)
    pl = emul.EmulPlan(lib, g)
    rng = np.random.default_rng(5)
    mk = lambda: np.ascontiguousarray(rng.standard_normal(g.kshape) + 1j * rng.standard_normal(g.kshape))
    n = 3
    start, d1, d2 = [mk() for _ in range(n)], [mk() for _ in range(n)], [mk() for _ in range(n)]
    coeff = [0.0, 0.004, 0.3]        # None-branch, Taylor branch, exp branch
    dt = 0.05
    for vo in (1, 2):
        k2p = g.k2() ** vo
        Zmax = [c * k2p.max() * dt for c in coeff]
        if vo == 1:
            assert Zmax[1] < 0.5 and Zmax[2] > 0.5
        for kind, fn in ((0, orc.euler), (1, orc.etd1), (2, orc.etd2rk1), (3, orc.etd2rk2)):
            out = pl.stage(kind, start, d1, d2, coeff, vo, dt)
            for c in range(n):
                ref = np.empty_like(start[c])
                IF = -(coeff[c] * k2p)
                if kind == 0:
                    fn(start[c], ref, d1[c], dt)
                elif kind == 1:
                    fn(start[c], ref, d1[c], IF, dt)
                else:
                    fn(start[c], ref, d1[c], d2[c], IF, dt)
                assert rel(out[c], ref) < 1e-15, (kind, c, vo)


@pytest.mark.parametrize("shape", [(16, 32), (8, 16, 16)])
def test_emulated_retained_only_sweep_equals_full_sweep(lib, shape):
    """DDL_STAGE_RETAINED_ONLY visits only the modes inside the dealias mask; with operands that
    vanish outside it the result must equal the full sweep bit for bit."""
    g = orc.Grid(shape, None)
    pl = emul.EmulPlan(lib, g)
    rng = np.random.default_rng(9)
    keep = ~g.dealias_mask()
    mk = lambda: np.ascontiguousarray((rng.standard_normal(g.kshape) + 1j * rng.standard_normal(g.kshape)) * keep)
    start, d1, d2 = [mk() for _ in range(2)], [mk() for _ in range(2)], [mk() for _ in range(2)]
    for kind in (1, 2, 3):
        full = pl.stage(kind, start, d1, d2, [0.01, 0.2], 1, 0.05, flags=0)
        part = pl.stage(kind, start, d1, d2, [0.01, 0.2], 1, 0.05, flags=1)
        for a, b in zip(full, part):
            assert np.array_equal(a * keep, b * keep)
            assert np.all(a[~keep] == 0)


FUSE_SHAPES = [("IncompressibleMHD", (8, 16, 16)), ("BoussinesqHydro", (16, 8, 16)), ("IncompressibleHydro", (16, 32)),
               ("IncompressibleMHD", (16, 16))]


def _fuse_case(lib, physics, shape):
    g = orc.Grid(shape, None)
    pl = emul.EmulPlan(lib, g)
    kw = {"direction": "y" if len(shape) == 2 else "z"} if physics == "BoussinesqHydro" else {}
    P = orc.PHYSICS[physics](shape, None, "2/3 cython", **kw)
    state = list(orc.synthetic_ic(P, 3).kvector())
    y = list(orc.synthetic_ic(P, 4).kvector())
    other = list(orc.synthetic_ic(P, 6).kvector())
    params = {"boussinesq_direction": "y" if len(shape) == 2 else "z"}
    coeff = ([0.0, 0.01, 0.3, 0.3, 0.02, 0.0])[:len(state)]
    k, _ = pl.rhs(physics, params, state, flags=1)
    return pl, params, state, y, other, coeff, list(k)


@pytest.mark.parametrize("physics,shape", FUSE_SHAPES)
@pytest.mark.parametrize("first,last", [(1, 0), (0, 0), (0, 1)])
def test_emulated_fused_rk4_stage_equals_rhs_then_stage(lib, physics, shape, first, last):
    """ddl_rhs_stage (derivative consumed in registers) == ddl_rhs followed by ddl_rk4_stage."""
    pl, params, state, y, total, coeff, k = _fuse_case(lib, physics, shape)
    ref_out, ref_total = pl.rk4_stage(y, k, total, coeff, 1, 3.0, 0.05, first, last)
    r = pl.rhs_stage(physics, params, state, 4, y, coeff, 1, 0.05, total=total, wdiv=3.0, first=first, last=last)
    assert rel(r["out"], ref_out) < 1e-15
    if not last:
        assert rel(r["total"], ref_total) < 1e-15


@pytest.mark.parametrize("physics,shape", FUSE_SHAPES)
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_emulated_fused_etd_stages_equal_rhs_then_stage(lib, physics, shape, kind):
    """Fused EULER / ETD1 (derivative also stored, as RK2's first stage needs) and ETD2RK1 / ETD2RK2 (derivative
    of this RHS is the second one, the first is read) == ddl_rhs followed by ddl_stage (retained-only sweep)."""
    pl, params, state, y, d_first, coeff, k = _fuse_case(lib, physics, shape)
    if kind in (0, 1):
        ref = np.stack(pl.stage(kind, y, k, None, coeff, 1, 0.05, flags=1))
        r = pl.rhs_stage(physics, params, state, kind, y, coeff, 1, 0.05, want_k=True)
        assert rel(r["k"], np.stack(k)) < 1e-15
    else:
        ref = np.stack(pl.stage(kind, y, d_first, k, coeff, 1, 0.05, flags=1))
        r = pl.rhs_stage(physics, params, state, kind, y, coeff, 1, 0.05, deriv1=d_first)
    assert rel(r["out"], ref) < 1e-15


@pytest.mark.parametrize("physics,shape", FUSE_SHAPES)
def test_emulated_fused_cn_step_equals_rhs_then_cn(lib, physics, shape):
    pl, params, state, y, _, coeff, k = _fuse_case(lib, physics, shape)
    ref = pl.cn_step(state, k, coeff, 1, 0.05)                 # CN updates the state it evaluated the RHS on
    r = pl.rhs_stage(physics, params, state, 5, state, coeff, 1, 0.05)
    assert rel(r["out"], ref) < 1e-15


@pytest.mark.parametrize("shape", [(16, 16, 16), (16, 32), (8, 32, 16)])
@pytest.mark.parametrize("physics", ["IncompressibleHydro", "BoussinesqHydro", "IncompressibleMHD"])
def test_emulated_advective_policies_match_reference_form_for_compressive_states(lib, physics, shape):
    """include/ddl.h DDL_*_ADV: for a state with div u, div B != 0 the conservative pipeline is O(1) away from
    the reference's advective forms (physics.py:197-228); the advective-form policies reproduce them."""
    kw = dict(direction="y") if (physics == "BoussinesqHydro" and len(shape) == 2) else {}
    P = orc.PHYSICS[physics](shape, **kw)
    if physics != "IncompressibleHydro":
        P.parameters.update(dict(g=1.3, beta=0.7, rho0=0.6))
    d = P.create_fields(0.)
    rng = np.random.default_rng(1)
    for _, _, c in d.components():
        c["xspace"] = rng.standard_normal(P.g.shape)
        c.require_space("kspace")
    y = d.kvector()
    k = P.create_fields(0.)
    P.RHS(d, k)
    params = dict(P.parameters)
    if kw:
        params["boussinesq_direction"] = "y"
    pl = emul.EmulPlan(lib, P.g)
    dk, _ = pl.rhs(physics, params, y, adv=True)
    assert rel(dk, k.kvector()) < 1e-14
    dk0, _ = pl.rhs(physics, params, y, adv=False)
    assert rel(dk0, k.kvector()) > 0.1


@pytest.mark.parametrize("physics,shape,cfg", [("IncompressibleMHD", (16, 16, 16), 5), ("BoussinesqHydro", (8, 16, 32), 4),
                                               ("IncompressibleHydro", (12, 20, 24), 3)])
def test_emulated_plane_chunked_rhs_equals_the_default(lib, physics, shape, cfg):
    """ddl_set_option("rhs_plane_chunk", n): the same RHS with y_inv -> x -> y_fwd over chunks of n z-planes and reused
    chunk-sized arrays (an opt-in L2-residency experiment); chunk sizes that do and do not divide nz."""
    kw = {"direction": "z"} if physics == "BoussinesqHydro" else {}
    Po = orc.PHYSICS[physics](shape, None, "2/3 cython", **kw)
    Po.parameters.update(dict(nu=1e-3))
    y0 = orc.synthetic_ic(Po, cfg).kvector()
    pl = emul.EmulPlan(lib, orc.Grid(shape))
    params = dict(Po.parameters)
    if kw:
        params["boussinesq_direction"] = "z"
    ref, _ = pl.rhs(physics, params, list(y0))
    try:
        for n in (1, 3, shape[0]):
            assert lib.ddl_set_option(b"rhs_plane_chunk", n) == 0
            d, _ = pl.rhs(physics, params, list(y0))
            assert np.array_equal(d, ref), n
    finally:
        lib.ddl_set_option(b"rhs_plane_chunk", 0)


@pytest.mark.parametrize("nd", [2, 3])
def test_emulated_array_factor_kernels_match_reference_cython(lib, nd):
    """ddl_step_array = euler / etd1 / etd2rk1 / etd2rk2 with the reference's own signature (integrating factor as an array)
    against the DIRECT outputs of the reference's Cython kernels (tests/golden/stage_kernels.npz): Z == 0, |Z| < 0.5 and the
    closed-form branch all occur, 2-D takes f0 from the series (forward_step_cy_2d.pyx:55)."""
    import ctypes as C
    z = np.load(os.path.join(GOLDEN, "stage_kernels.npz"))
    p = "k%dd_" % nd
    s, d1, d2, IF = (np.ascontiguousarray(z[p + n]) for n in ("start", "d1", "d2", "IF"))
    dt = float(z[p + "dt"])
    lib.ddl_step_array.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_double, C.c_void_p]
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    for kind, name, two in ((0, "euler", False), (1, "etd1", False), (2, "etd2rk1", True), (3, "etd2rk2", True)):
        o = np.full_like(s, np.nan)
        rc = lib.ddl_step_array(kind, nd, s.size, ptr(s), ptr(o), ptr(d1), ptr(d2) if two else None, ptr(IF), dt, None)
        assert rc == 0, lib.ddl_last_error()
        assert rel(o, z[p + name]) < 1e-15, name
    # no factor at all: the Euler forms of the integrators' `integrating_factor is None` branches
    o = np.empty_like(s)
    assert lib.ddl_step_array(1, nd, s.size, ptr(s), ptr(o), ptr(d1), None, None, dt, None) == 0
    assert np.array_equal(o, s + dt * d1)
    assert lib.ddl_step_array(3, nd, s.size, ptr(s), ptr(o), ptr(d1), ptr(d2), None, dt, None) == 0
    assert np.array_equal(o, s + dt * d2)
    assert lib.ddl_step_array(2, nd, s.size, ptr(s), ptr(o), ptr(d1), ptr(d2), None, dt, None) == 0
    assert rel(o, s + dt / 2. * (d2 - d1)) < 1e-16
    assert lib.ddl_step_array(7, nd, s.size, ptr(s), ptr(o), ptr(d1), None, None, dt, None) != 0


@pytest.mark.parametrize("nd", [2, 3])
@pytest.mark.parametrize("branch", ["row", "dense"])
def test_emulated_dealias_array_matches_reference_cython(lib, nd, branch):
    """ddl_dealias_array = dealias_23 with the reference's own signature against the DIRECT outputs of its Cython kernels
    (tests/golden/dealias_kernels.npz): ky per row and ky dense (the shearing box's branch)."""
    import ctypes as C
    z = np.load(os.path.join(GOLDEN, "dealias_kernels.npz"))
    p = "d%d_" % nd
    data = np.ascontiguousarray(z[p + "data"]).copy()
    kx = np.ascontiguousarray(z[p + "kx"].ravel())
    ky = np.ascontiguousarray(z[p + ("kydense" if branch == "dense" else "ky")]).reshape(-1)
    kz = np.ascontiguousarray(z[p + "kz"].ravel()) if nd == 3 else None
    kny = np.ascontiguousarray(z[p + "kny"])
    shape = np.array(data.shape, dtype=np.int64)
    lib.ddl_dealias_array.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = lib.ddl_dealias_array(nd, ptr(shape), ptr(data), ptr(kx), ptr(ky), ptr(kz), 1 if branch == "dense" else 0, ptr(kny), None)
    assert rc == 0, lib.ddl_last_error()
    want = z[p + branch]
    assert np.array_equal(data, want)
    assert 0 < (want == 0).sum() < want.size


@pytest.mark.parametrize("shape", [(16, 32), (8, 16, 16), (12, 20, 24)])
def test_emulated_outside_mask_count(lib, shape):
    """ddl_reduce_outside_mask: non-zero entries outside the dealias mask, per array (the check behind verify_clean)."""
    import ctypes as C
    g = orc.Grid(shape)
    pl = emul.EmulPlan(lib, g)
    rng = np.random.default_rng(2)
    mask = g.dealias_mask()
    clean = rng.standard_normal(g.kshape) + 1j * rng.standard_normal(g.kshape)
    clean[mask] = 0.0
    junk = clean.copy()
    idx = np.argwhere(mask)[::7]
    junk[tuple(idx.T)] = 1e-300                       # tiny, but not zero
    nan = clean.copy()
    nan[tuple(np.argwhere(mask)[0])] = np.nan
    arrays = [np.ascontiguousarray(a) for a in (clean, junk, nan, clean * 2)]
    out = np.full(8, -1.0)
    lib.ddl_reduce_outside_mask.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    ptrs = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    assert lib.ddl_reduce_outside_mask(pl.plan, len(arrays), ptrs, out.ctypes.data_as(C.c_void_p), None) == 0, lib.ddl_last_error()
    assert out[:4].tolist() == [0.0, float(len(idx)), 1.0, 0.0]
    assert all(np.array_equal(a, b) for a, b in zip(arrays, (clean, junk, nan, clean * 2)) if not np.isnan(a).any())
