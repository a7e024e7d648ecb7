"""Run the reference's own code (oracle/_ref through oracle/ref_run.py) on a given state in a child process
and return what it produced.  TEST INFRASTRUCTURE: the child is the checker, never the thing measured."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_RUN = os.path.join(ROOT, "oracle", "ref_run.py")


def ref_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", ".built"))


def start_reference(tmpdir, physics, shape, y0, integ, steps, dt, params, threads=None, direction=None, length=None):
    """Launch the child; returns a handle for finish_reference (so the device run can overlap it)."""
    tmpdir = str(tmpdir)
    y0file, out = os.path.join(tmpdir, "y0.npy"), os.path.join(tmpdir, "ref")
    np.save(y0file, np.ascontiguousarray(y0))
    cmd = [sys.executable, REF_RUN, "--physics", physics, "--shape"] + [str(s) for s in shape] + [
        "--integ", integ, "--steps", str(steps), "--dt", repr(float(dt)), "--y0", y0file, "--out", out,
        "--threads", str(threads or os.cpu_count() or 1)]
    if length:
        cmd += ["--length"] + [repr(float(v)) for v in length]
    if direction:
        cmd += ["--direction", direction]
    for k, v in (params or {}).items():
        cmd += ["--param", "%s=%r" % (k, float(v))]
    env = dict(os.environ)
    env.pop("DEDALUS_DDL_LIB", None)
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env), out


def finish_reference(handle, timeout=1500):
    proc, out = handle
    so, se = proc.communicate(timeout=timeout)
    if proc.returncode != 0:
        raise RuntimeError("reference child failed:\n" + se[-3000:])
    meta = json.load(open(os.path.join(out, "meta.json")))
    return np.load(os.path.join(out, "y1.npy")), meta


def run_reference(tmpdir, *a, **kw):
    return finish_reference(start_reference(tmpdir, *a, **kw))
