"""CPU: the reduction entry points (ddl_reduce_invariants, ddl_reduce_max_square, ddl_rhs_capture_max)
through the g++ host-emulation build of the kernel bodies, against the oracle's restatement of
dedalus/analysis/volume_average.py and fields.py:153-157.  The -m gpu tests repeat it on the device
(tests/test_gpu_widen.py)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "host"))
import dedalus_oracle as orc
import emul

# include/ddl.h DDL_INV_*
EKIN, E2, DIV_SUM, MAG_DIV_SUM, ENST, CUR2, HEL_KIN, HEL_CROSS = range(8)
DIV_RE, DIV_IM, MDIV_RE, MDIV_IM, CENK_NUM, CENK_DEN, MSQ = 8, 9, 10, 11, 12, 13, 14
HEL_MAG, GRAD2_T = 20, 21


@pytest.fixture(scope="module")
def lib():
    return emul.load()


def random_state(P, seed, solenoidal=True, dealiased=True):
    """Hermitian-consistent random state: real noise -> forward (-> mask) [-> div_free]."""
    d = P.create_fields(0.)
    rng = np.random.default_rng(seed)
    for _, _, c in d.components():
        c["xspace"] = rng.standard_normal(P.g.shape)
        c.require_space("kspace")
        if not dealiased:
            # forward() dealiases; put energy back outside the mask so that the full sweep differs
            m = P.g.dealias_mask()
            c.kdata[m] = (rng.standard_normal(int(m.sum())) + 1j * rng.standard_normal(int(m.sum()))) * 1e-3
    if solenoidal:
        for name, f in d:
            if f.ncomp > 1:
                f.div_free()
    return d


def close(a, b, rtol=1e-12, atol=1e-13):
    return abs(a - b) <= atol + rtol * abs(b)


def check_invariants(inv, ref, ncomp):
    assert close(inv[EKIN], ref["ekin"])
    assert close(inv[DIV_SUM], ref["div_sum"], atol=1e-11)
    assert close(inv[ENST], ref["enstrophy"])
    assert close(inv[DIV_RE], ref["divergence"].real, atol=1e-11) and close(inv[DIV_IM], ref["divergence"].imag, atol=1e-11)
    assert close(inv[CENK_NUM], ref["cenk_num"]) and close(inv[CENK_DEN], ref["cenk_den"])
    for c in range(ncomp):
        assert close(inv[MSQ + c], ref["msq"][c])
    for key, idx in (("e2", E2), ("mag_div_sum", MAG_DIV_SUM), ("current2", CUR2), ("hel_kin", HEL_KIN),
                     ("hel_cross", HEL_CROSS), ("hel_mag", HEL_MAG), ("grad2_T", GRAD2_T)):
        if key in ref:
            assert close(inv[idx], ref[key], atol=1e-11), key
    if "mag_div" in ref:
        assert close(inv[MDIV_RE], ref["mag_div"].real, atol=1e-11) and close(inv[MDIV_IM], ref["mag_div"].imag, atol=1e-11)


CASES = [("IncompressibleHydro", (16, 32)), ("BoussinesqHydro", (16, 32)), ("IncompressibleMHD", (32, 16)),
         ("IncompressibleHydro", (8, 16, 32)), ("BoussinesqHydro", (16, 16, 16)), ("IncompressibleMHD", (16, 8, 32))]


@pytest.mark.parametrize("physics,shape", CASES)
@pytest.mark.parametrize("solenoidal", [True, False])
def test_invariants_match_reference_tasks(lib, physics, shape, solenoidal):
    P = orc.PHYSICS[physics](shape)
    d = random_state(P, 11, solenoidal=solenoidal)
    pl = emul.EmulPlan(lib, P.g)
    y = d.kvector()
    ref = orc.invariants(d)
    full = pl.invariants(physics, y, flags=0)
    compact = pl.invariants(physics, y, flags=1)        # DDL_STAGE_RETAINED_ONLY: the state is dealiased
    check_invariants(full, ref, len(y))
    check_invariants(compact, ref, len(y))
    if not solenoidal:
        assert ref["div_sum"] > 1e-3                     # the divergence entries are really exercised


@pytest.mark.parametrize("physics,shape", [("IncompressibleMHD", (16, 16)), ("IncompressibleMHD", (8, 16, 16))])
def test_invariants_full_sweep_sees_modes_outside_the_mask(lib, physics, shape):
    P = orc.PHYSICS[physics](shape)
    d = random_state(P, 5, solenoidal=False, dealiased=False)
    pl = emul.EmulPlan(lib, P.g)
    y = d.kvector()
    ref = orc.invariants(d)
    check_invariants(pl.invariants(physics, y, flags=0), ref, len(y))
    assert not close(pl.invariants(physics, y, flags=1)[EKIN], ref["ekin"], rtol=1e-9)


def test_invariants_known_answers(lib):
    """Taylor-Green vortex u = (sin x cos y, -cos x sin y): <u^2>/2 = 1/4, enstrophy <w^2>/2 = 1/2, zero
    divergence (set in x-space: the reference's spectral table init_cond.py:33-51 carries Nyquist junk,
    SURVEY F7); ABC flow (curl u = u): helicity = 2 * energy = 2 * enstrophy."""
    P = orc.IncompressibleHydro((32, 32))
    d = P.create_fields(0.)
    yy, xx = np.meshgrid(*(np.arange(32) * 2 * np.pi / 32,) * 2, indexing="ij")
    d["u"][0]["xspace"] = np.sin(xx) * np.cos(yy)
    d["u"][1]["xspace"] = -np.cos(xx) * np.sin(yy)
    pl = emul.EmulPlan(lib, P.g)
    inv = pl.invariants("IncompressibleHydro", d.kvector(), flags=0)
    assert close(inv[EKIN], 0.25) and close(inv[ENST], 0.5) and inv[DIV_SUM] < 1e-14
    P3 = orc.IncompressibleHydro((16, 16, 16))
    d3 = P3.create_fields(0.)
    n = 16
    z, y, x = np.meshgrid(*(np.arange(n) * 2 * np.pi / n,) * 3, indexing="ij")
    d3["u"][0]["xspace"] = np.sin(z) + np.cos(y)          # ABC flow, A = B = C = 1: curl u = u
    d3["u"][1]["xspace"] = np.sin(x) + np.cos(z)
    d3["u"][2]["xspace"] = np.sin(y) + np.cos(x)
    pl3 = emul.EmulPlan(lib, P3.g)
    inv = pl3.invariants("IncompressibleHydro", d3.kvector(), flags=1)
    assert close(inv[EKIN], 1.5) and close(inv[HEL_KIN], 3.0) and close(inv[ENST], 1.5)


@pytest.mark.parametrize("physics,shape", CASES)
def test_max_square_matches_reference(lib, physics, shape):
    P = orc.PHYSICS[physics](shape)
    d = random_state(P, 21)
    pl = emul.EmulPlan(lib, P.g)
    y = d.kvector()
    params = dict(P.parameters, boussinesq_direction="y" if len(shape) == 2 else "z")
    out, _ = pl.max_square(physics, params, y)
    ref = orc.max_squares(d)
    assert close(out[0], ref[0], rtol=1e-13)
    assert close(out[1], ref[1], rtol=1e-13)


def test_max_square_dealiases_in_place_like_the_reference(lib):
    """fields.py:153-157 reads c['xspace'], i.e. backward(): the spectrum is masked in place
    (representations.py:347-357) and the maximum is that of the dealiased field."""
    physics, shape = "IncompressibleMHD", (16, 16, 16)
    P = orc.PHYSICS[physics](shape)
    d = random_state(P, 2, dealiased=False)
    pl = emul.EmulPlan(lib, P.g)
    y = d.kvector()
    out, y_after = pl.max_square(physics, P.parameters, y, flags=2)      # DDL_RHS_DEALIAS_STATE
    ref = orc.max_squares(d)                                             # dealiases d's components in place
    assert close(out[0], ref[0], rtol=1e-13) and close(out[1], ref[1], rtol=1e-13)
    y_masked = y.copy()
    y_masked[:, P.g.dealias_mask()] = 0.0
    assert np.array_equal(y_after, y_masked)
    out2, y_same = pl.max_square(physics, P.parameters, y, flags=0)      # values identical, state untouched
    assert np.array_equal(out2, out) and np.array_equal(y_same, y)


@pytest.mark.parametrize("physics,shape", [("IncompressibleMHD", (16, 32)), ("IncompressibleHydro", (16, 16, 16)),
                                           ("IncompressibleMHD", (8, 16, 32))])
def test_capture_inside_the_rhs_changes_nothing_else(lib, physics, shape):
    P = orc.PHYSICS[physics](shape)
    d = random_state(P, 31)
    pl = emul.EmulPlan(lib, P.g)
    y = d.kvector()
    deriv0, _ = pl.rhs(physics, P.parameters, y)
    cap = np.zeros(2)
    pl.capture_max(cap)
    deriv1, _ = pl.rhs(physics, P.parameters, y)
    pl.capture_max(None)
    assert np.array_equal(deriv0, deriv1)
    ref = orc.max_squares(d)
    assert close(cap[0], ref[0], rtol=1e-13) and close(cap[1], ref[1], rtol=1e-13)
    # accumulating semantics: a second state only raises the maxima
    cap2 = cap.copy()
    pl.capture_max(cap2)
    pl.rhs(physics, P.parameters, 0.5 * y)
    pl.capture_max(None)
    assert np.array_equal(cap2, cap)
    frozen = cap2.copy()
    pl.rhs(physics, P.parameters, 3.0 * y)               # capture off: untouched
    assert np.array_equal(cap2, frozen)


def test_slab_invariants_sum_to_the_global_ones(lib):
    """Each rank reduces its own ky slab; the sums over ranks are the global invariants (block and
    cyclic ownership)."""
    physics, shape = "IncompressibleMHD", (16, 16, 16)
    P = orc.PHYSICS[physics](shape)
    d = random_state(P, 41, solenoidal=False)
    y = d.kvector()
    ref = emul.EmulPlan(lib, P.g).invariants(physics, y, flags=1)
    additive = [EKIN, E2, DIV_SUM, MAG_DIV_SUM, ENST, CUR2, HEL_KIN, HEL_CROSS, DIV_RE, DIV_IM, HEL_MAG] + list(range(MSQ, MSQ + 6))
    ny = shape[1]
    for layout in (0, 1):
        for nranks in (2, 4):
            tot = np.zeros(24)
            for r in range(nranks):
                rows = np.arange(r * ny // nranks, (r + 1) * ny // nranks) if layout == 0 else np.arange(r, ny, nranks)
                pl = emul.EmulPlan(lib, P.g, nranks=nranks, rank=r, layout=layout)
                tot += pl.invariants(physics, [np.ascontiguousarray(c[rows]) for c in y], flags=1)
            for i in additive:
                assert close(tot[i], ref[i], atol=1e-11), (layout, nranks, i)
