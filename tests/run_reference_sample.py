"""Child process of tests/test_reference_samples.py: execute one of the REFERENCE's own sample scripts, unmodified, against the
drop-in package (the script's `from dedalus.mods import *` resolves to dedalus-1.0_b200/dedalus), stopping after a few
iterations.  usage: run_reference_sample.py <script> <max_iterations> <workdir>"""
import os
import runpy
import sys

script, max_iter, workdir = sys.argv[1], int(sys.argv[2]), sys.argv[3]
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import conftest  # noqa: F401,E402   (paths; the host-emulation harness when DDL_TEST_HOST_EMUL=1)
import dedalus.time_stepping.time_step as ts  # noqa: E402

_ok = ts.TimeStepBase.ok


def ok(self):
    return self.iteration < max_iter and _ok.fget(self)


ts.TimeStepBase.ok = property(ok)
os.chdir(workdir)
src = open(script).read()
try:
    compile(src, script, "exec")
    ns = runpy.run_path(script, run_name="__main__")
except SyntaxError:
    # a Python-2 script: the mechanical edits of oracle/build_ref.py (print statements, xrange, ...), in memory only
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    import build_ref
    ns = {"__name__": "__main__", "__file__": script}
    exec(compile(build_ref.transliterate("sample", src), script, "exec"), ns)
import numpy as np  # noqa: E402
data, ti = ns["data"], ns["ti"]
state = np.stack([c["kspace"].cpu().numpy() for _, _, c in data.components()])
assert np.isfinite(state).all() and np.linalg.norm(state) > 0
print("SAMPLE_OK iterations=%d time=%.6e norm=%.12e" % (ti.iteration, ti.time, np.linalg.norm(state)))
