"""CPU: the kernels' index logic under AddressSanitizer + UBSan.  The host-emulation build executes the very same kernel bodies
(global-memory indexing, shared-memory tile indexing -- the tile is a malloc'ed block there --, loop bounds); instrumented with
-fsanitize=address,bounds,shift,... (dedalus-1.0_b200/build.py build_emul(sanitize=True)) every out-of-bounds read or write of a
field buffer, a workspace or a tile aborts the child process.  This is the CPU-side stand-in for compute-sanitizer, which needs
the device: an out-of-bounds access that happens to land in mapped memory passes every numerical test unnoticed.
Test infrastructure only; the instrumented library lives in tests/host/_build_asan/ and is never shipped."""
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _libasan():
    for pat in ("/usr/lib/x86_64-linux-gnu/libasan.so.*", "/usr/lib/gcc/x86_64-linux-gnu/*/libasan.so"):
        hits = sorted(glob.glob(pat))
        if hits:
            return os.path.realpath(hits[0])
    return None


ASAN = _libasan()
pytestmark = pytest.mark.skipif(ASAN is None, reason="no libasan runtime in this image")


@pytest.fixture(scope="module", autouse=True)
def instrumented_lib():
    """Built once here, before any child (xdist workers would otherwise race to build it)."""
    sys.path.insert(0, os.path.join(ROOT, "dedalus-1.0_b200"))
    import build as ddl_build
    return ddl_build.build_emul(os.path.join(ROOT, "tests", "host", "_build_asan"), sanitize=True)


def _env():
    env = dict(os.environ, LD_PRELOAD=ASAN, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:abort_on_error=0",
               UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1", DDL_TEST_HOST_EMUL_ASAN="1", OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")
    env.pop("DEDALUS_DDL_LIB", None)
    return env


def _pytest(files, extra_env=None, marker=None):
    try:
        import xdist  # noqa: F401
        par = ["-n", str(min(6, os.cpu_count() or 1))]
    except ImportError:
        par = []
    env = _env()
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider"] + par + (["-m", marker] if marker else []) + \
          [os.path.join(ROOT, "tests", f) for f in files]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=2400)
    tail = "\n".join(r.stdout.strip().splitlines()[-25:])
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail and "AddressSanitizer" not in r.stdout and "runtime error" not in r.stdout, tail


def test_instrumented_build_reports_an_overrun():
    """The harness itself: a deliberately short output buffer handed to ddl_deriv must abort the child with an ASan report."""
    code = (
        "import sys, os, ctypes as C, numpy as np\n"
        "sys.path.insert(0, os.path.join(%r, 'tests', 'host')); sys.path.insert(0, os.path.join(%r, 'oracle'))\n"
        "import emul, dedalus_oracle as orc\n"
        "lib = emul.load(); g = orc.Grid((8, 8)); pl = emul.EmulPlan(lib, g)\n"
        "k = np.zeros(g.kshape, dtype=np.complex128); o = np.zeros(k.size - 3, dtype=np.complex128)\n"
        "lib.ddl_deriv(pl.plan, k.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), 0, None)\n"
        "print('survived')\n") % (ROOT, ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=_env(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode != 0 and "AddressSanitizer" in r.stdout and "heap-buffer-overflow" in r.stdout, r.stdout[-2000:]
    assert "survived" not in r.stdout


def test_kernel_bodies_under_asan():
    """tests/test_host_emulation.py + the any-length transforms (C ABI through ctypes, numpy buffers) + the reference's six
    sample scripts (their children inherit the instrumented library and the preloaded runtime through the environment)."""
    _pytest(["test_host_emulation.py", "test_mixed_radix.py", "test_host_reductions.py", "test_reference_samples.py"], marker="not gpu")


def test_drop_in_package_under_asan():
    """The `-m gpu` test files through the drop-in package (torch CPU buffers), every physics / integrator / grid they cover."""
    _pytest(["test_gpu_parity.py", "test_gpu_widen.py", "test_mixed_radix.py", "test_gpu_shear.py", "test_gpu_restart.py",
             "test_gpu_analysis.py", "test_gpu_known_answers.py", "test_gpu_errors.py"], {"DDL_TEST_HOST_EMUL": "1"}, marker="gpu")


@pytest.mark.parametrize("world,layout,worker", [(2, "block", "gpu_slab_worker.py"), (4, "cyclic", "slab_unfused_worker.py")])
def test_slab_pipeline_under_asan(world, layout, worker, tmp_path):
    """The slab-decomposed pipeline (pack / unpack index tables of the exchanges, per-rank pencil passes) over gloo."""
    import json
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "res.json")
    env = dict(_env(), DDL_TEST_HOST_EMUL="1", DEDALUS_KY_LAYOUT=layout, DEDALUS_SLAB_EXCHANGE="collective")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", worker), out]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert r.returncode == 0 and "AddressSanitizer" not in r.stdout and "runtime error" not in r.stdout, r.stdout[-3000:]
    assert len(json.load(open(out))) >= 5
