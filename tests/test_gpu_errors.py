"""Error behaviour at the drop-in boundary: the exceptions the reference raises, at the same sites (ValueError
representations.py:105-107,245,339,351; KeyError :156; NotImplementedError :305-310,382; ValueError physics.py:494), plus the
library's own refusals surfacing as exceptions instead of wrong answers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _native_lib_loaded():
    from conftest import native_lib_expected
    native_lib_expected()
    yield


L2, L3 = (2 * np.pi,) * 2, (2 * np.pi,) * 3


def test_representation_argument_errors():
    from dedalus.data_objects.api import FourierRepresentation
    with pytest.raises(ValueError, match="2 or 3 dimensions"):
        FourierRepresentation(None, (16,), (1.0,))
    with pytest.raises(ValueError, match="2 or 3 dimensions"):
        FourierRepresentation(None, (8, 8, 8, 8), (1.0,) * 4)
    with pytest.raises(ValueError, match="same dimensions"):
        FourierRepresentation(None, (16, 16), L3)


def test_space_state_machine_errors():
    from dedalus.data_objects.api import FourierRepresentation
    c = FourierRepresentation(None, (16, 16), L2)
    with pytest.raises(KeyError):
        c["pspace"] = 0.
    with pytest.raises(ValueError):
        c.require_space("pspace")
    with pytest.raises(ValueError, match="Forward transform cannot be called from kspace"):
        c.forward()
    c["xspace"]
    with pytest.raises(ValueError, match="Backward transform cannot be called from xspace"):
        c.backward()


def test_configuration_errors():
    from dedalus.config import decfg
    from dedalus.data_objects.api import FourierRepresentation
    try:
        decfg.set("FFT", "method", "abacus")
        with pytest.raises(NotImplementedError, match="FFT method"):
            FourierRepresentation(None, (16, 16), L2)
        decfg.set("FFT", "method", "cuda")
        decfg.set("FFT", "dealiasing", "5/7")
        with pytest.raises(NotImplementedError, match="dealiasing"):
            FourierRepresentation(None, (16, 16), L2)
    finally:
        decfg.set("FFT", "method", "cuda")
        decfg.set("FFT", "dealiasing", "2/3 cython")


def test_physics_errors():
    from dedalus.mods import IncompressibleHydro, IncompressibleMHD, FourierRepresentation
    P = IncompressibleHydro((16, 16), FourierRepresentation)
    with pytest.raises(KeyError):
        P["no_such_parameter"]
    P.parameters["shear_rate"] = 0.5
    with pytest.raises(ValueError, match="shearing representation"):
        P.RHS(P.create_fields(0.), P.create_fields(0.))
    M = IncompressibleMHD((16, 16, 16), FourierRepresentation)
    data = M.create_fields(0.)
    with pytest.raises(ValueError):
        data.add_field("u", "VectorField")                    # state_data.py:139-140


def test_library_refusals_surface_as_exceptions():
    """Lengths the kernels cannot factor, and the error text of the C ABI, reach the caller as exceptions."""
    import dedalus._lib as L
    from dedalus.data_objects.api import FourierRepresentation
    with pytest.raises(L.DDLError, match="prime factors"):
        FourierRepresentation(None, (2 * 67, 16), L2)
    with pytest.raises(L.DDLError, match="unsupported"):
        FourierRepresentation(None, (4096, 16), L2)
    # and nothing of that sticks: the next plan is fine
    c = FourierRepresentation(None, (12, 20), L2)
    c["xspace"] = 1.0
    assert abs(complex(c["kspace"][0, 0].item()) - 1.0) < 1e-15
