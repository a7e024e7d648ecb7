"""Parity at BASELINE 3-D sizes against the reference's own code ("256^3 x 5 steps once per
round"): the device RK4 (Python API -> C ABI -> sm_100a kernels) vs the reference's physics.py RHS +
representations.py transforms + verbatim Cython euler / etd1 kernels, driven through the restated RK4 data
flow of time_step.py:426-483 by oracle/ref_run.py in a child process (FFTs threaded, the reference's code
untouched).  Tolerance: relative L2 <= 1e-10 on the spectral state (north star), plus the reference's own
ekin / emag / divergence_sum tasks.  The child runs on the host cores while the device run takes milliseconds."""
import numpy as np
import pytest

from devutil import rel, dev_physics, oracle_physics, set_state, get_state
from refglue import ref_available, start_reference, finish_reference

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_available(), reason="oracle/_ref is not built")]
TOL = 1e-10


@pytest.fixture(scope="module", autouse=True)
def _native_lib_loaded():
    from conftest import native_lib_expected
    native_lib_expected()
    yield


def _device_run(physics, shape, params, y0, steps, dt):
    import torch
    import dedalus.time_stepping.api as tapi
    import dedalus.analysis.volume_average as va
    P = dev_physics(physics, shape, None, params)
    data = P.create_fields(0.)
    set_state(data, y0)
    ti = tapi.RK4(P)
    for _ in range(steps):
        ti.do_advance(data, dt)
    torch.cuda.synchronize()
    inv = {"ekin": va.ekin(data), "divergence_sum": va.divergence_sum(data), "time": data.time}
    if physics == "IncompressibleMHD":
        inv["emag"], inv["mag_div_sum"] = va.emag(data), va.mag_div_sum(data)
    y1 = get_state(data)
    del data, ti, P
    torch.cuda.empty_cache()
    return y1, inv


@pytest.mark.parametrize("physics,n,steps,params", [
    ("IncompressibleMHD", 256, 3, dict(nu=1e-3, eta=1e-3)),            # the headline physics, 1/8 of its size
    ("BoussinesqHydro", 256, 3, dict(nu=1e-3, kappa=1e-3)),            # BASELINE config 4's physics
    ("IncompressibleHydro", 256, 5, dict(nu=1e-3)),                    # BASELINE config 3 at full size
    ("IncompressibleMHD", 32, 4, dict(nu=1e-3, eta=1e-3)),             # the same harness at a size the emulated run takes too
])
def test_rk4_vs_reference_code(tmp_path, physics, n, steps, params):
    import dedalus_oracle as orc
    shape = (n, n, n)
    Po = oracle_physics(physics, shape, None, params)
    do = orc.synthetic_ic(Po, 5)
    y0 = do.kvector().copy()
    umax = np.sqrt(max(orc.max_squares(do)))
    del do, Po
    dt = 0.2 * (2 * np.pi / n) / umax
    child = start_reference(tmp_path, physics, shape, y0, "RK4", steps, dt, params,
                            direction="z" if physics == "BoussinesqHydro" else None)
    try:
        y1, inv = _device_run(physics, shape, params, y0, steps, dt)
    finally:
        ref, meta = finish_reference(child)
    err = rel(y1, ref)
    print("%s %d^3 x %d RK4 steps vs the reference's code: rel L2 = %.3e (reference child: %.1f s on %d threads)"
          % (physics, n, steps, err, meta["seconds"], meta["threads"]))
    assert err < TOL
    assert abs(inv["time"] - meta["time"]) < 1e-14
    assert abs(inv["ekin"] - meta["ekin"]) < 1e-12 * max(1.0, abs(meta["ekin"]))
    # sums of |round-off| over all modes: both at round-off level, not equal to each other
    assert inv["divergence_sum"] / n ** 3 < 1e-12 and meta["divergence_sum"] / n ** 3 < 1e-12
    if physics == "IncompressibleMHD":
        assert abs(inv["emag"] - meta["emag"]) < 1e-12 * max(1.0, abs(meta["emag"]))
        assert inv["mag_div_sum"] / n ** 3 < 1e-12 and meta["mag_div_sum"] / n ** 3 < 1e-12
    # per-component agreement, so that a small component cannot hide behind a large one
    for j in range(y1.shape[0]):
        assert rel(y1[j], ref[j]) < TOL


@pytest.mark.parametrize("config,physics,shape,integ,steps,dt,params", [
    ("config 1: Taylor-Green 128^2 RK2mid", "IncompressibleHydro", (128, 128), "RK2mid", 20, 5e-3, dict(nu=1e-2)),
    ("config 2: Orszag-Tang 512^2 RK4", "IncompressibleMHD", (512, 512), "RK4", 5, 2e-3, dict(nu=1e-3, eta=1e-3)),
])
def test_baseline_2d_configs_vs_reference_code(tmp_path, config, physics, shape, integ, steps, dt, params):
    """BASELINE configs 1 and 2 at their full sizes against the reference's own code: config 1 through the reference's RK2mid class
    itself (time_step.py:224-309), with its Taylor-Green field exactly as init_cond.py:33-51 writes it (Nyquist-row entries
    included); config 2 through its RHS and Cython kernels under the restated RK4 glue."""
    import torch
    import dedalus_oracle as orc
    import dedalus.time_stepping.api as tapi
    Po = oracle_physics(physics, shape, None, params)
    do = Po.create_fields(0.)
    if physics == "IncompressibleHydro":
        orc.taylor_green(do)
    else:
        orc.orszag_tang(do)
    y0 = do.kvector().copy()
    child = start_reference(tmp_path, physics, shape, y0, integ, steps, dt, params, threads=4)
    try:
        P = dev_physics(physics, shape, None, params)
        data = P.create_fields(0.)
        set_state(data, y0)
        ti = getattr(tapi, integ)(P)
        for _ in range(steps):
            ti.do_advance(data, dt)
        torch.cuda.synchronize() if torch.cuda.is_available() else None
        y1 = get_state(data)
    finally:
        ref, meta = finish_reference(child)
    err = rel(y1, ref)
    print("%s vs the reference's code: rel L2 = %.3e" % (config, err))
    assert err < TOL
    assert abs(data.time - meta["time"]) < 1e-13
