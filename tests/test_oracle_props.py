"""CPU: what pins the two integrators the reference cannot run as shipped (RK4,
CrankNicholsonVisc: time_step.py:209,214,449,495).  Their restatement in
oracle/dedalus_oracle.py is "parity unpinned" by the reference; it is pinned here by
  (i)   order of accuracy: RK4 dt^4 with every integrating factor None, dt^2 with them active
        (forward_step is a first-order ETD step, time_step.py:187-221), CN dt^1,
  (ii)  agreement with the reference-pinned RK2mid as dt -> 0,
  (iii) the exact Taylor-Green decay exp(-2 nu t) (init_cond.py:33-51).
"""
import numpy as np
import pytest

import dedalus_oracle as orc


def _run(integ, nu, dt, nsteps, seed=3, shape=(16, 16)):
    P = orc.IncompressibleHydro(shape)
    P.parameters["nu"] = nu
    d = orc.synthetic_ic(P, seed)
    ti = getattr(orc, integ)(P)
    for _ in range(nsteps):
        ti.do_advance(d, dt)
    return d.kvector()


def _orders(integ, nu, T=0.4, ns=(4, 8, 16), nref=256):
    ref = _run("RK4", nu, T / nref, nref)
    errs = [np.linalg.norm(_run(integ, nu, T / n, n) - ref) / np.linalg.norm(ref) for n in ns]
    return [np.log2(errs[i] / errs[i + 1]) for i in range(len(errs) - 1)], errs


def test_rk4_is_fourth_order_without_integrating_factors():
    orders, errs = _orders("RK4", 0.0)
    assert all(3.7 < o < 4.3 for o in orders), (orders, errs)


def test_rk4_is_second_order_with_integrating_factors():
    orders, errs = _orders("RK4", 0.05)
    assert all(1.8 < o < 2.3 for o in orders), (orders, errs)


def test_crank_nicholson_is_first_order():
    orders, errs = _orders("CrankNicholsonVisc", 0.05, ns=(16, 32, 64))
    assert all(0.8 < o < 1.25 for o in orders), (orders, errs)


@pytest.mark.parametrize("integ", ["RK4", "CrankNicholsonVisc"])
def test_restated_integrators_converge_to_rk2mid(integ):
    T, n = 0.2, 512
    a = _run(integ, 0.02, T / n, n)
    b = _run("RK2mid", 0.02, T / n, n)
    tol = 1e-6 if integ == "RK4" else 2e-3          # CN is first order
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < tol


@pytest.mark.parametrize("integ", ["RK4", "RK2mid"])
def test_taylor_green_decays_exactly(integ):
    nu, dt, nsteps = 0.1, 0.01, 40
    P = orc.IncompressibleHydro((32, 32))
    P.parameters["nu"] = nu
    d = P.create_fields(0.0)
    orc.taylor_green(d)
    ti = getattr(orc, integ)(P)
    for _ in range(nsteps):
        ti.do_advance(d, dt)
    exact = -1j / 4.0 * np.exp(-2 * nu * dt * nsteps)
    assert abs(d["u"][0].kdata[1, 1] - exact) < 1e-14
