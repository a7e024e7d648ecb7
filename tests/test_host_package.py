"""CPU: the drop-in Python package end to end (physics classes, integrators, diagnostics, CFL) with the
kernels' bodies executed by the g++ host-emulation build -- the `-m gpu` test files re-run in a child
process under tests/conftest.py's opt-in harness (DDL_TEST_HOST_EMUL=1; grids above 64^3 and the tests that
need a real device are skipped there).  This is how the host logic is checked in the GPU-less build
container; the product package itself has no CPU path (tests/test_abi.py asserts that)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_test_files_pass_under_host_emulation():
    env = dict(os.environ, DDL_TEST_HOST_EMUL="1")
    env.pop("DEDALUS_DDL_LIB", None)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                        os.path.join(ROOT, "tests", "test_gpu_parity.py"), os.path.join(ROOT, "tests", "test_gpu_widen.py")],
                       cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    tail = "\n".join(r.stdout.strip().splitlines()[-15:])
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
