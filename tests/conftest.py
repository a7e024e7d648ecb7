import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dedalus-1.0_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# ---------------------------------------------------------------------------------------------
# Host-emulation harness (TEST INFRASTRUCTURE, opt-in): DDL_TEST_HOST_EMUL=1 runs the `-m gpu`
# tests in the GPU-less build container.  The drop-in Python package is pointed at the g++
# host-emulation build of the same kernel sources (tests/host/_build/libddl_emul.so, never
# shipped, never loaded by the package on its own) and its two CUDA-specific helpers are patched
# IN THIS TEST PROCESS ONLY to CPU tensors / a null stream.  It checks the host logic (physics,
# integrators, diagnostics, slab bookkeeping) and the kernels' index logic before GPU time is
# spent; it proves nothing about the device build, which is what the real `-m gpu` run is for.
# The product itself has no CPU path: without this variable `plan.device()` raises when CUDA is absent.
# ---------------------------------------------------------------------------------------------
HOST_EMUL = os.environ.get("DDL_TEST_HOST_EMUL") == "1"
HOST_EMUL_MAX_POINTS = int(os.environ.get("DDL_TEST_HOST_EMUL_MAX_POINTS", str(64 ** 3)))


def _enable_host_emulation():
    sys.path.insert(0, os.path.join(ROOT, "tests", "host"))
    import build as ddl_build                   # dedalus-1.0_b200/build.py
    if os.environ.get("DDL_TEST_HOST_EMUL_ASAN") == "1":      # tests/test_host_sanitizer.py: instrumented build, libasan preloaded
        os.environ["DEDALUS_DDL_LIB"] = ddl_build.build_emul(os.path.join(ROOT, "tests", "host", "_build_asan"), sanitize=True)
    else:
        os.environ["DEDALUS_DDL_LIB"] = ddl_build.build_emul(os.path.join(ROOT, "tests", "host", "_build"))
    import torch
    import dedalus.data_objects.plan as plan_mod
    plan_mod.device = lambda: torch.device("cpu")
    plan_mod.current_stream = lambda: None
    import dedalus._lib as L
    assert b"host emulation" in L.lib.ddl_version()


if HOST_EMUL:
    _enable_host_emulation()


def native_lib_expected():
    """What the -m gpu tests assert about the loaded library."""
    import dedalus._lib as L
    if HOST_EMUL:
        assert b"host emulation" in L.lib.ddl_version()
        return
    import torch
    assert torch.cuda.is_available()
    assert b"sm_100a" in L.lib.ddl_version()


def _points(item):
    """Largest grid among the parameters of a test item (shape tuples or an edge length `n`)."""
    best = 0
    params = getattr(getattr(item, "callspec", None), "params", {})
    for k, v in params.items():
        if isinstance(v, (tuple, list)) and v and all(isinstance(i, int) for i in v):
            n = 1
            for i in v:
                n *= i
            best = max(best, n)
        elif k == "n" and isinstance(v, int):
            best = max(best, v ** 3)
    return best


# Order of the -m gpu run (the driver uses -x): what has already passed on a B200 first, then the files added after the
# round's GPU budget was spent (verified through the host emulation and, for the kernels, by tests/native/devcheck.cu), so that
# a device-only surprise in new code cannot hide the state of everything before it.
_FILE_ORDER = ["test_gpu_parity", "test_gpu_slab", "test_native_abi", "test_gpu_widen", "test_mixed_radix", "test_gpu_restart",
               "test_gpu_shear", "test_gpu_analysis", "test_gpu_known_answers", "test_gpu_errors"]
_NEW_GOLDENS = ("10x30", "50_rk2trap", "18x24", "48x2x48", "9x15x14", "12x20x24", "nodealias")


def _order_key(item):
    name = os.path.basename(str(item.fspath))[:-3]
    rank = _FILE_ORDER.index(name) if name in _FILE_ORDER else -1
    if name == "test_gpu_parity" and any(g in item.nodeid for g in _NEW_GOLDENS):
        rank = _FILE_ORDER.index("test_mixed_radix")        # grids that are not powers of two: with their own file
    return rank


def pytest_collection_modifyitems(config, items):
    items.sort(key=_order_key)          # stable: the order inside a file is kept
    if not HOST_EMUL:
        return
    big = pytest.mark.skip(reason="host emulation: grid too large for one CPU thread")
    needs_device = pytest.mark.skip(reason="host emulation: needs the CUDA device (graphs, peer memory, NCCL)")
    for item in items:
        if _points(item) > HOST_EMUL_MAX_POINTS or "full_size" in item.name or "512" in item.name or "256cubed" in item.name \
                or "128cubed" in item.name or "async_staging" in item.name:
            item.add_marker(big)
        if "cuda_graph" in item.name or "test_gpu_slab" in item.nodeid:
            item.add_marker(needs_device)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
