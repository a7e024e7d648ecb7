"""Snapshot / restart (reference: dedalus/time_stepping/time_step.py:112-151, dedalus/utils/restart.py:33-99): a run
continued from its snapshot equals the uninterrupted run bit for bit; forcing functions come back by name from the
sidecar source file; the integrator's statistics and the integrating-factor coefficients survive the pickle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _native_lib_loaded():
    from conftest import native_lib_expected
    native_lib_expected()
    yield


def kvec(d):
    return np.stack([c["kspace"].cpu().numpy().copy() for _, _, c in d.components()])


def spin(x, t):
    """A forcing function that lives in a real source file, so that the snapshot can carry its source."""
    return x


def velocity_forcing(data, i):
    return 0.05 * data["u"][i]["kspace"].clone() if i == 0 else 0.0 * data["u"][i]["kspace"].clone()


def noise_state(P, shape, seed):
    import torch
    data = P.create_fields(0.)
    rng = np.random.default_rng(seed)
    for _, f in data:
        for _, c in f:
            c["xspace"] = torch.from_numpy(rng.standard_normal(shape))
            c["kspace"]
        if f.ncomp > 1:
            f.div_free()
    return data


@pytest.mark.parametrize("physics,shape,integ", [("IncompressibleMHD", (16, 16, 16), "RK2mid"), ("IncompressibleMHD", (16, 16, 16), "RK4"),
                                                 ("BoussinesqHydro", (32, 16), "RK2trap"), ("IncompressibleHydro", (10, 30), "CrankNicholsonVisc")])
def test_restart_continues_bit_for_bit(tmp_path, monkeypatch, physics, shape, integ):
    import dedalus.time_stepping.api as tapi
    from dedalus.utils.api import restart
    from devutil import dev_physics
    monkeypatch.chdir(tmp_path)
    P = dev_physics(physics, shape, None, dict(nu=0.01))
    data = noise_state(P, shape, 4)
    ti = getattr(tapi, integ)(P)
    for _ in range(3):
        ti.do_advance(data, 5e-3)
    ti.snapshot(data)
    assert os.path.exists("snap_00000/dedalus_obj_0000.cpkl") and os.path.exists("snap_00000/forcing_functions.py")
    for _ in range(2):
        ti.do_advance(data, 5e-3)
    want = kvec(data)
    R2, d2, t2 = restart("snap_00000")
    assert t2.iteration == 3 and abs(t2.time - 0.015) < 1e-15 and abs(d2.time - 0.015) < 1e-15
    assert type(R2) is type(P) and R2.parameters["nu"] == 0.01
    for _ in range(2):
        t2.do_advance(d2, 5e-3)
    got = kvec(d2)
    # the first step after a restart re-derives the integrating-factor bookkeeping on the unfused path: equal to round-off,
    # and bit-identical for the integrators whose first step is always unfused
    assert np.linalg.norm(got - want) <= 1e-14 * np.linalg.norm(want)
    assert t2.iteration == 5 and abs(t2.time - ti.time) < 1e-15


def test_restart_reattaches_forcing_functions(tmp_path, monkeypatch):
    from dedalus.mods import IncompressibleHydro, FourierRepresentation, RK2mid
    from dedalus.utils.api import restart
    monkeypatch.chdir(tmp_path)
    P = IncompressibleHydro((16, 16), FourierRepresentation)
    P.parameters["nu"] = 0.02
    P.set_velocity_forcing(velocity_forcing)
    data = noise_state(P, (16, 16), 9)
    ti = RK2mid(P)
    ti.do_advance(data, 1e-2)
    ti.snapshot(data)
    assert "def velocity_forcing" in open("snap_00000/forcing_functions.py").read()
    ti.do_advance(data, 1e-2)
    R2, d2, t2 = restart("snap_00000")
    assert R2.forcing_functions["VelocityForcing"].__name__ == "velocity_forcing"
    t2.do_advance(d2, 1e-2)
    assert np.linalg.norm(kvec(d2) - kvec(data)) <= 1e-14 * np.linalg.norm(kvec(data))


def test_snapshot_in_x_space_restores_the_space(tmp_path, monkeypatch):
    """The 'space' attribute of every component is saved and restored (fields.py:118-125, restart.py:88-97)."""
    from dedalus.mods import IncompressibleHydro, FourierRepresentation, RK2mid
    from dedalus.utils.api import restart
    monkeypatch.chdir(tmp_path)
    P = IncompressibleHydro((16, 16), FourierRepresentation)
    data = noise_state(P, (16, 16), 2)
    x = data["u"]["x"]["xspace"].cpu().numpy().copy()
    ti = RK2mid(P)
    ti.snapshot(data)
    R2, d2, t2 = restart("snap_00000")
    assert d2["u"]["x"]._curr_space == "xspace" and d2["u"]["y"]._curr_space == "kspace"
    assert np.array_equal(d2["u"]["x"]["xspace"].cpu().numpy(), x)


def test_load_all_reads_a_snapshot(tmp_path, monkeypatch):
    """load_all (parallelism.py:54-100): one component of a snapshot as a single array, axes 0 and 1 swapped in 3-D."""
    from dedalus.mods import IncompressibleMHD, FourierRepresentation, RK2mid
    from dedalus.utils.api import load_all
    monkeypatch.chdir(tmp_path)
    P = IncompressibleMHD((8, 16, 16), FourierRepresentation)
    data = noise_state(P, (8, 16, 16), 3)
    RK2mid(P).snapshot(data)
    arr, space = load_all("B/1", "snap_00000")
    want = data["B"][1]["kspace"].cpu().numpy()
    assert space == "kspace" and np.array_equal(arr, np.transpose(want, [1, 0, 2]))
    arr2, _ = load_all("/fields/u/0", "snap_00000")
    assert arr2.shape == (8, 16, 9)


def test_restart_in_a_shearing_box(tmp_path, monkeypatch):
    """The drifting wavenumbers are rebuilt from the restored time (state_data.py:115-121 on load)."""
    import torch
    from dedalus.mods import IncompressibleHydro, FourierShearRepresentation, RK2mid
    from dedalus.utils.api import restart
    monkeypatch.chdir(tmp_path)
    P = IncompressibleHydro((16, 32), FourierShearRepresentation)
    P.parameters.update(dict(nu=0.01, shear_rate=1.5))
    data = P.create_fields(0.4)
    rng = np.random.default_rng(6)
    for _, c in data["u"]:
        c["xspace"] = torch.from_numpy(rng.standard_normal((16, 32)))
        c["kspace"]
    data["u"].div_free()
    ti = RK2mid(P)
    for _ in range(2):
        ti.do_advance(data, 1e-2)
    ti.snapshot(data)
    for _ in range(2):
        ti.do_advance(data, 1e-2)
    R2, d2, t2 = restart("snap_00000")
    assert abs(d2.time - 0.42) < 1e-14 and not d2["u"][0]._static_k
    for _ in range(2):
        t2.do_advance(d2, 1e-2)
    assert np.array_equal(d2["u"][0].k["y"].cpu().numpy(), data["u"][0].k["y"].cpu().numpy())
    assert np.linalg.norm(kvec(d2) - kvec(data)) <= 1e-14 * np.linalg.norm(kvec(data))


def test_hdf5_branch_of_snapshot_and_restart_with_an_h5py_stand_in(tmp_path, monkeypatch):
    """This image has no h5py, so the HDF5 branch of snapshot() / restart() / load_all() / identify_version() -- the
    reference's format (time_step.py:141-151, fields.py:118-125, restart.py:33-110, parallelism.py:54-100) -- is executed
    against tests/h5shim.py, a stand-in for the h5py calls it makes: same layout (/time, hg_version, /fields/<name>/<comp>,
    attribute 'space'), same continuation as the .npy fallback.  The HDF5 byte format itself stays unverified."""
    import sys
    import h5shim
    import dedalus.time_stepping.api as tapi
    from dedalus.utils.api import restart
    from dedalus.utils.restart import identify_version
    from dedalus.utils.parallelism import load_all
    from devutil import dev_physics
    monkeypatch.chdir(tmp_path)
    monkeypatch.setitem(sys.modules, "h5py", h5shim)
    shape = (16, 16, 16)
    P = dev_physics("IncompressibleMHD", shape, None, dict(nu=0.01, eta=0.02))
    data = noise_state(P, shape, 6)
    ti = tapi.RK2mid(P)
    for _ in range(2):
        ti.do_advance(data, 5e-3)
    data["B"]["y"]["xspace"]                      # one component is saved in x-space: its 'space' attribute must bring it back
    ti.snapshot(data)
    assert os.path.exists("snap_00000/data.cpu0000") and not os.path.exists("snap_00000/fields.cpu0000.json")
    with h5shim.File("snap_00000/data.cpu0000", "r") as f:
        assert abs(float(np.asarray(f["time"][...])) - ti.time) < 1e-15 and f.attrs["hg_version"]
        assert f["/fields/u"].attrs["type"] == "VectorField" and f["/fields/u"].attrs["representation"] == "FourierRepresentation"
        assert f["/fields/u/0"].attrs["space"] == "kspace" and f["/fields/B/1"].attrs["space"] == "xspace"
        assert f["/fields/u/0"].shape == (16, 16, 9) and f["/fields/B/1"].shape == (16, 16, 16)
    assert identify_version("snap_00000") == f.attrs["hg_version"]
    arr, space = load_all("u/0", "snap_00000")
    assert space == "kspace" and arr.shape == (16, 16, 9)
    for _ in range(2):
        ti.do_advance(data, 5e-3)
    ref = kvec(data)
    RHS2, data2, ti2 = restart("snap_00000")
    assert data2["B"]["y"]._curr_space == "xspace"
    for _ in range(2):
        ti2.do_advance(data2, 5e-3)
    assert np.abs(kvec(data2) - ref).max() < 1e-15 and ti2.iteration == ti.iteration
