"""The C ABI driven from plain C++ (tests/native/devcheck.cu): no Python, no torch between the caller and
include/ddl.h -- what a Cython / ctypes binding inside the reference would do.  CPU: the same program against the
host-emulation build (checks the checker and the ABI conventions: plan creation from wavenumber / mask arrays,
workspace sizing, pointer lists).  GPU: the device library -- reductions vs a host loop, CFL capture vs explicit
inverse transforms, every x-pass variant vs the generic tile kernels."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "native"))


def _run(exe, args, env=None):
    r = subprocess.run([exe] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "all ok" in r.stdout, r.stdout[-4000:]
    return r.stdout


@pytest.mark.parametrize("n", [16, 32])
def test_devcheck_host_emulation(n):
    import build_devcheck
    out = _run(build_devcheck.emul(), [n])
    assert "host emulation" in out and out.count(" ok") >= 12 and "FAIL" not in out


def test_devcheck_device_binary_builds():
    """nvcc cross-compiles the checker against the CUDA library (nothing is run without a GPU)."""
    import build_devcheck
    exe = build_devcheck.device()
    assert os.path.exists(exe)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [64, 128, 256])
def test_devcheck_device(n):
    import build_devcheck
    out = _run(build_devcheck.device(), [n], env=build_devcheck.cuda_env())
    assert "sm_100a" in out and "FAIL" not in out


@pytest.mark.gpu
def test_devcheck_device_timing_mode():
    """The native RK4 step (4 x ddl_rhs_stage) and the per-kernel profile: the 20-second measurement loop for kernel work."""
    import build_devcheck
    out = _run(build_devcheck.device(), [0, 128, "reps=2"], env=build_devcheck.cuda_env())
    assert "RK4 step, fused stages" in out and "x_fused" in out and "assemble_stage" in out
