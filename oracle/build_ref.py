#!/usr/bin/env python
"""Materialise a runnable copy of the reference's hot-path files under oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing on the product path may import oracle/.

The reference (/root/reference, jsoishi/dedalus-1.0) is Python 2 + Cython and cannot be
imported by the Python 3.12 interpreter of this image.  This script applies the purely
*mechanical* Py2->Py3 edits listed in SURVEY.md section 8(c) to a scratch copy of the
hot-path files, compiles the reference's four Cython kernels **verbatim**, and leaves the
result in ``oracle/_ref/`` (git-ignored, but shipped to the GPU box by gpurun).  No
arithmetic is changed: the numpy-FFT backend (``FFT.method = numpy``,
dedalus/data_objects/representations.py:303-307,327-333) is the reference's own.

Used for
  * generating the golden vectors in tests/golden/ (tests/golden/make_golden.py),
  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg (reference CPU path
    timed on the GPU host's cores).

Re-run:  python oracle/build_ref.py        (needs /root/reference; a no-op message otherwise)
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DEDALUS_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

PY_FILES = [
    "dedalus/__init__.py",
    "dedalus/config.py",
    "dedalus/funcs.py",
    "dedalus/utils/__init__.py",
    "dedalus/utils/parallelism.py",
    "dedalus/utils/timer.py",
    "dedalus/utils/function_count.py",
    "dedalus/utils/logger.py",
    "dedalus/utils/misc_numeric.py",
    "dedalus/data_objects/__init__.py",
    "dedalus/data_objects/api.py",
    "dedalus/data_objects/aux_equation.py",
    "dedalus/data_objects/fields.py",
    "dedalus/data_objects/state_data.py",
    "dedalus/data_objects/representations.py",
    "dedalus/physics/__init__.py",
    "dedalus/physics/api.py",
    "dedalus/physics/physics.py",
    "dedalus/time_stepping/__init__.py",
    "dedalus/time_stepping/api.py",
    "dedalus/time_stepping/time_step.py",
    "dedalus/analysis/__init__.py",
    "dedalus/analysis/volume_average.py",
    "dedalus/init_cond/__init__.py",
    "dedalus/init_cond/init_cond.py",
    "dedalus/init_cond/turb_spectra.py",
]

PYX_FILES = {
    "forward_step_cy_2d": "dedalus/time_stepping/forward_step_cy_2d.pyx",
    "forward_step_cy_3d": "dedalus/time_stepping/forward_step_cy_3d.pyx",
    "dealias_cy_2d": "dedalus/data_objects/dealias_cy_2d.pyx",
    "dealias_cy_3d": "dedalus/data_objects/dealias_cy_3d.pyx",
}


def _convert_prints(src: str) -> str:
    """print statement -> print() call, joining parenthesis-continued lines."""
    lines = src.split("\n")
    out = []
    i = 0
    while i < len(lines):
        line = lines[i]
        m = re.match(r"^(\s*)print(\s+(.*))?$", line)
        if m and not line.lstrip().startswith("print("):
            indent, rest = m.group(1), (m.group(3) or "")
            # continuation: unbalanced parentheses
            while rest.count("(") > rest.count(")") and i + 1 < len(lines):
                i += 1
                rest += " " + lines[i].strip()
            out.append("%sprint(%s)" % (indent, rest))
        else:
            out.append(line)
        i += 1
    return "\n".join(out)


GENERIC = [
    (r"\bxrange\b", "range"),
    (r"\.iteritems\(\)", ".items()"),
    (r"\.itervalues\(\)", ".values()"),
    (r"(\w+(?:\.\w+)*)\.has_key\(([^)]*)\)", r"(\2 in \1)"),
    (r"\.func_name\b", ".__name__"),
    (r"^import ConfigParser", "import configparser as ConfigParser"),
    (r"^import cPickle", "import pickle as cPickle"),
    (r"np\.asfarray\(([^)]*)\)", r"np.asarray(\1, dtype=float)"),
]

SPECIFIC = {
    "dedalus/data_objects/state_data.py": [
        (r"^from fields import", "from .fields import"),
        (r"^import h5py", "h5py = None"),
        (r"self\._field_classes\.keys\(\)\[0\]", "list(self._field_classes.keys())[0]"),
        (r"self\.fields\.keys\(\)\[0\]", "list(self.fields.keys())[0]"),
        (r"field_keys = zip\((.*)\)$", r"field_keys = list(zip(\1))"),
    ],
    "dedalus/time_stepping/api.py": [
        (r"^from time_step import", "from .time_step import"),
    ],
    "dedalus/time_stepping/time_step.py": [
        (r"^import h5py", "h5py = None"),
        (r"from forward_step_cy_(\dd) import", r"from forward_step_cy_\1 import"),
    ],
    "dedalus/data_objects/representations.py": [
        (r"^from dedalus\.utils\.fftw import fftw", "fftw = None"),
        # integer division (Py2 '/' on ints)
        (r"self\.global_shape\['kspace'\]\[-1\] / 2 \+ 1", "self.global_shape['kspace'][-1] // 2 + 1"),
        (r"ki\[ksize / 2\]", "ki[ksize // 2]"),
        (r"self\.local_shape\['kspace'\]\[1\] / 2", "self.local_shape['kspace'][1] // 2"),
        (r"np\.array\(plane_data\.shape\) / 2", "np.array(plane_data.shape) // 2"),
        # list-of-slices indexing -> tuple
        (r"sli = \[slice\(i\) for i in self\.data\.shape\]", "sli = tuple(slice(i) for i in self.data.shape)"),
        (r"index = zip\(\*test\.nonzero\(\)\)", "index = list(zip(*test.nonzero()))"),
        (r"refgrid\[\[slice\(i\) for i in np\.asarray\(self\.local_shape\['xspace'\], dtype=float\)\]\]",
         "refgrid[tuple(slice(float(i)) for i in self.local_shape['xspace'])]"),
        # np.ogrid[...] returned a list in the numpy the reference was written for and returns a tuple now
        (r"^(\s*)grid\[0\] \+= self\.offset\['xspace'\]", r"\1grid = list(grid) if open else grid\n\1grid[0] += self.offset['xspace']"),
    ],
    "dedalus/physics/physics.py": [
        (r"if self\.k2 == None:", "if self.k2 is None:"),
        (r"if self\.parameters\['Omega'\] == None:", "if self.parameters['Omega'] is None:"),
    ],
    "dedalus/funcs.py": [
        (r"^import inspect", "import inspect\nimport logging\nmylog = logging.getLogger('Dedalus')"),
    ],
    "dedalus/analysis/volume_average.py": [
        (r"if k == None:", "if k is None:"),
    ],
    "dedalus/init_cond/init_cond.py": [
        (r"kshape/2 \+ 1", "kshape//2 + 1"),
        (r"ux\.data\.shape\[1\]/2 \+ 1", "ux.data.shape[1]//2 + 1"),
    ],
    "dedalus/utils/parallelism.py": [],
}


def transliterate(rel: str, src: str) -> str:
    src = _convert_prints(src)
    for pat, rep in GENERIC:
        src = re.sub(pat, rep, src, flags=re.M)
    for pat, rep in SPECIFIC.get(rel, []):
        src, n = re.subn(pat, rep, src, flags=re.M)
    return src


def build(force: bool = False) -> bool:
    if not os.path.isdir(REF):
        print("oracle/build_ref.py: %s not present; keeping prebuilt oracle/_ref as is" % REF)
        return os.path.isdir(OUT)
    stamp = os.path.join(OUT, ".built")
    if os.path.exists(stamp) and not force:
        return True
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    for rel in PY_FILES:
        with open(os.path.join(REF, rel)) as f:
            src = f.read()
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w") as f:
            f.write(transliterate(rel, src))
    # F12: the reference needs a built __hg_version__ module
    with open(os.path.join(OUT, "dedalus", "__hg_version__.py"), "w") as f:
        f.write("hg_version = 'oracle'\n")
    # utils/api.py pulls restart.py (h5py) and metrics; the hot path only needs these names
    with open(os.path.join(OUT, "dedalus", "utils", "api.py"), "w") as f:
        f.write("from .function_count import counts\nfrom .timer import Timer, timer\n"
                "from .parallelism import com_sys, swap_indices\n")
    # analysis/api.py pulls matplotlib; the oracle only needs volume_average
    # Cython kernels, verbatim
    cy = os.path.join(OUT, "cy")
    os.makedirs(cy)
    for name, rel in PYX_FILES.items():
        shutil.copyfile(os.path.join(REF, rel), os.path.join(cy, name + ".pyx"))
    with open(os.path.join(cy, "setup.py"), "w") as f:
        f.write(
            "import numpy\nfrom setuptools import setup, Extension\n"
            "from Cython.Build import cythonize\n"
            "exts = [Extension(n, [n + '.pyx'], include_dirs=[numpy.get_include()],\n"
            "                  extra_compile_args=['-O3', '-w']) for n in %r]\n"
            "setup(ext_modules=cythonize(exts, language_level=2, quiet=True))\n" % (sorted(PYX_FILES),)
        )
    subprocess.check_call([sys.executable, "setup.py", "-q", "build_ext", "--inplace"], cwd=cy,
                          stdout=subprocess.DEVNULL)
    # drop the copied .pyx / generated .c: only the compiled kernels stay
    for fn in os.listdir(cy):
        if fn.endswith((".pyx", ".c")) or fn == "setup.py":
            os.remove(os.path.join(cy, fn))
    shutil.rmtree(os.path.join(cy, "build"), ignore_errors=True)
    open(stamp, "w").write("ok\n")
    return True


def import_ref():
    """Put oracle/_ref on sys.path and return the reference modules (numpy FFT backend)."""
    if not os.path.exists(os.path.join(OUT, ".built")):
        if not build():
            raise RuntimeError("oracle/_ref is not built and %s is absent" % REF)
    for p in (OUT, os.path.join(OUT, "cy")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from dedalus.config import decfg
    decfg.set("FFT", "method", "numpy")
    decfg.set("utils", "loglevel", "warning")
    import dedalus.data_objects.api as data_api
    import dedalus.physics.api as physics_api
    import dedalus.time_stepping.time_step as time_step
    return decfg, data_api, physics_api, time_step


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable")
