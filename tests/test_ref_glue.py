"""The restated RK4 / CrankNicholsonVisc of the oracle against the REFERENCE'S OWN CODE under the glue of
SURVEY.md 8(c): oracle/ref_run.py drives the reference's physics.py RHS, its representations.py (numpy FFT
backend) and its verbatim-compiled Cython euler / etd1 kernels through the RK4 data flow of
time_step.py:426-483.  This pins the oracle's RK4 (the integrator of the headline benchmark, which the
reference cannot run as shipped) to the reference's code instead of to a restatement of it.
RK2mid through the same child is the reference's integrator class, unmodified: the harness checks itself."""
import numpy as np
import pytest

import dedalus_oracle as orc
from devutil import rel, oracle_physics
from refglue import ref_available, run_reference

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref is not built (needs /root/reference)")

CASES = [
    ("IncompressibleHydro", (24, 24), dict(nu=1e-3), 5e-3),
    ("IncompressibleHydro", (16, 16, 16), dict(nu=0.0), 5e-3),                      # inviscid: euler path (time_step.py:290-291)
    ("IncompressibleMHD", (32, 32), dict(nu=1e-3, eta=2e-3), 4e-3),
    ("IncompressibleMHD", (16, 24, 32), dict(nu=1e-3, eta=1e-3), 4e-3),
    ("IncompressibleMHD", (16, 16, 16), dict(nu=2.0, eta=3.0), 2e-2),               # stiff: |Z| > 0.5, exp branch (forward_step_cy_3d.pyx:57-59)
    ("BoussinesqHydro", (16, 16, 16), dict(nu=1e-3, kappa=2e-3), 5e-3),
    ("BoussinesqHydro", (24, 16), dict(nu=0.5, kappa=0.0), 1e-2),                   # mixed: u with IF, T without
]


@pytest.mark.parametrize("physics,shape,params,dt", CASES)
def test_oracle_rk4_equals_reference_code_under_restated_glue(tmp_path, physics, shape, params, dt):
    Po = oracle_physics(physics, shape, None, params)
    do = orc.synthetic_ic(Po, 3)
    y0 = do.kvector().copy()
    direction = ("y" if len(shape) == 2 else "z") if physics == "BoussinesqHydro" else None
    y1, meta = run_reference(tmp_path, physics, shape, y0, "RK4", 3, dt, params, threads=1, direction=direction)
    ti = orc.RK4(Po)
    for _ in range(3):
        ti.do_advance(do, dt)
    assert rel(do.kvector(), y1) < 1e-13
    assert abs(do.time - meta["time"]) < 1e-14
    assert abs(orc.energy(do, "u") - meta["ekin"]) < 1e-13
    if physics == "IncompressibleMHD":
        assert abs(orc.energy(do, "B") - meta["emag"]) < 1e-13


def test_harness_reproduces_the_reference_rk2mid(tmp_path):
    """Same child, the reference's own RK2mid class: equals the oracle's RK2mid (which the goldens pin)."""
    physics, shape, params, dt = "IncompressibleMHD", (16, 16, 16), dict(nu=1e-2, eta=1e-2), 5e-3
    Po = oracle_physics(physics, shape, None, params)
    do = orc.synthetic_ic(Po, 4)
    y1, meta = run_reference(tmp_path, physics, shape, do.kvector().copy(), "RK2mid", 3, dt, params, threads=2)
    ti = orc.RK2mid(Po)
    for _ in range(3):
        ti.do_advance(do, dt)
    assert rel(do.kvector(), y1) < 1e-13


@pytest.mark.parametrize("physics,shape,params,dt", [CASES[0], CASES[3], CASES[5], CASES[6]])
def test_oracle_cn_equals_reference_code_under_restated_glue(tmp_path, physics, shape, params, dt):
    """CrankNicholsonVisc (time_step.py:486-506, not runnable as shipped) restated around the reference's own RHS."""
    Po = oracle_physics(physics, shape, None, params)
    do = orc.synthetic_ic(Po, 6)
    direction = ("y" if len(shape) == 2 else "z") if physics == "BoussinesqHydro" else None
    y1, meta = run_reference(tmp_path, physics, shape, do.kvector().copy(), "CrankNicholsonVisc", 3, dt, params, threads=1,
                             direction=direction)
    ti = orc.CrankNicholsonVisc(Po)
    for _ in range(3):
        ti.do_advance(do, dt)
    assert rel(do.kvector(), y1) < 1e-13
    assert abs(do.time - meta["time"]) < 1e-14


@pytest.mark.parametrize("physics,shape,integ,steps,dt,params", [
    ("IncompressibleHydro", (128, 128), "RK2mid", 20, 5e-3, dict(nu=1e-2)),          # BASELINE config 1 (the reference's own RK2mid class)
    ("IncompressibleMHD", (512, 512), "RK4", 2, 2e-3, dict(nu=1e-3, eta=1e-3)),      # BASELINE config 2 at full size
])
def test_oracle_equals_reference_code_on_the_2d_baseline_configs(tmp_path, physics, shape, integ, steps, dt, params):
    Po = oracle_physics(physics, shape, None, params)
    do = Po.create_fields(0.)
    (orc.taylor_green if physics == "IncompressibleHydro" else orc.orszag_tang)(do)
    y1, meta = run_reference(tmp_path, physics, shape, do.kvector().copy(), integ, steps, dt, params, threads=4)
    ti = orc.INTEGRATORS[integ](Po)
    for _ in range(steps):
        ti.do_advance(do, dt)
    assert rel(do.kvector(), y1) < 1e-13
    assert abs(orc.energy(do, "u") - meta["ekin"]) < 1e-13
