"""A caller who KEEPS the tensor `comp['kspace']` returned (in the reference it is the live numpy array) and writes to it
between steps -- forcing added into `uk` every step is the documented pattern -- must get the reference's answer: the
knowledge that selects the fused / retained-modes-only / conservative-form kernels is re-established whenever torch has
seen a write to the buffer since the last step (FourierRepresentation.refresh_escaped)."""
import numpy as np
import pytest

from devutil import rel, dev_physics, oracle_physics, set_state, get_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("integ", ["RK4", "RK2mid"])
@pytest.mark.parametrize("physics,shape,params", [("IncompressibleHydro", (32, 32, 32), dict(nu=0.02)),
                                                  ("IncompressibleMHD", (16, 32, 32), dict(nu=0.02, eta=0.03)),
                                                  ("IncompressibleHydro", (64, 64), dict(nu=0.01))])
def test_writes_through_a_kept_tensor_are_seen(physics, shape, params, integ):
    import torch
    import dedalus_oracle as orc
    import dedalus.time_stepping.api as tapi
    Po = oracle_physics(physics, shape, None, params)
    do = orc.synthetic_ic(Po, 31)
    P = dev_physics(physics, shape, None, params)
    data = P.create_fields(0.)
    set_state(data, do.kvector())
    ti, to = getattr(tapi, integ)(P), getattr(orc, integ)(Po)
    uk = data["u"]["x"]["kspace"]                    # kept for the whole run
    ux_o = do["u"][0]
    dt = 2e-3
    for _ in range(2):
        ti.do_advance(data, dt)
        to.do_advance(do, dt)
    comps = [c for _, _, c in data.components()]
    assert all(c._soln for c in comps) and comps[0]._escaped
    # a forcing increment that is NOT solenoidal and has content OUTSIDE the 2/3 mask, through the kept alias only
    rng = np.random.default_rng(3)
    inc = np.zeros(ux_o.kdata.shape, dtype=np.complex128)
    idx_in = (2, 3, 1) if len(shape) == 3 else (3, 2)
    idx_out = (shape[1] // 2 - 1, 2, 1) if len(shape) == 3 else (2, shape[0] // 2 - 1)
    inc[idx_in] = 0.2 - 0.1j
    inc[idx_out] = 0.05 + 0.02j
    for step in range(3):
        uk.add_(torch.from_numpy(inc).to(uk.device))
        ux_o.kdata[...] += inc
        ti.do_advance(data, dt)
        to.do_advance(do, dt)
        assert comps[0]._soln is False               # the compressive part was noticed: advective-form kernels
    assert rel(get_state(data), do.kvector()) < 1e-10
    assert abs(get_state(data)[0][idx_out]) > 0.01 or physics == "IncompressibleMHD"   # MHD dealiases its state, hydro keeps the junk


def test_touch_drops_everything_known_about_a_buffer():
    P = dev_physics("IncompressibleHydro", (16, 16, 16), None, dict(nu=0.01))
    data = P.create_fields(0.)
    c = data["u"]["x"]
    assert c._clean and c._sym and c._soln and not c._escaped
    c.touch()
    assert not c._clean and not c._sym and c._soln is None
