"""CPU: the C-ABI boundary.  The CUDA library built by __graft_entry__.build() loads without a GPU and
exports every function include/ddl.h declares (no compute call is made here); the Python binding lists
the same set; the host-emulation build used by the CPU tests exports them too; and the product package
refuses to work without a CUDA device instead of falling back to anything."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ddl.h")
LIB = os.path.join(ROOT, "dedalus-1.0_b200", "dedalus", "_lib", "libddl_b200.so")


def declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(ddl_[a-z0-9_]+)\s*\(", text))
    assert len(names) > 30
    return sorted(names)


def test_header_cites_the_reference_interfaces():
    text = open(HEADER).read()
    for needle in ("_fftw.pyx", "representations.py", "forward_step_cy_3d.pyx", "dealias_cy", "physics.py", "time_step.py"):
        assert needle in text, needle


def test_cuda_library_loads_and_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        import sys
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(LIB)
    missing = [n for n in declared() if not hasattr(lib, n)]
    assert not missing, missing
    lib.ddl_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.ddl_version()
    lib.ddl_launch_count.restype = ctypes.c_longlong
    assert lib.ddl_launch_count() == 0


def test_python_binding_lists_the_declared_symbols():
    src = open(os.path.join(ROOT, "dedalus-1.0_b200", "dedalus", "_lib", "__init__.py")).read()
    block = src[src.index("EXPORTS = ["):src.index("]", src.index("EXPORTS = ["))]
    listed = set(re.findall(r'"(ddl_[a-z0-9_]+)"', block))
    assert listed == set(declared()), (sorted(listed - set(declared())), sorted(set(declared()) - listed))


def test_host_emulation_build_exports_the_same_abi():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "host"))
    import emul
    lib = emul.load()
    missing = [n for n in declared() if not hasattr(lib, n)]
    assert not missing, missing
    lib.ddl_version.restype = ctypes.c_char_p
    assert b"host emulation" in lib.ddl_version()


def test_product_package_has_no_cpu_path():
    """Importing the package binds the CUDA library; building anything needs a CUDA device."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import torch\n"
            "from dedalus.mods import IncompressibleHydro, FourierRepresentation\n"
            "import dedalus._lib as L; assert b'sm_100a' in L.lib.ddl_version()\n"
            "try:\n    IncompressibleHydro((16, 16), FourierRepresentation).create_fields(0.)\n"
            "except RuntimeError as e:\n    print('REFUSED' if 'no CPU path' in str(e) else 'OTHER', e)\n"
            "else:\n    print('RAN')\n") % os.path.join(ROOT, "dedalus-1.0_b200")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=300)
    assert "REFUSED" in r.stdout, r.stdout[-2000:]


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use oracle/."""
    pkg = os.path.join(ROOT, "dedalus-1.0_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "dedalus_oracle" not in text and "oracle/" not in text and "import oracle" not in text, os.path.join(dirpath, f)
