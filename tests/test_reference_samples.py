"""CPU (build container only: needs /root/reference): the reference's OWN sample scripts -- the files under
/root/reference/samples that parse as Python 3 -- executed UNMODIFIED against the drop-in package, kernels through the host
emulation: `from dedalus.mods import *`, physics / representation / integrator construction, initial conditions, AnalysisSet
with VolumeAverage / TrackMode / PowerSpectrum / Snapshot tasks, the CFL-less advance loop, finalize.  A few iterations each.
What they exercise beyond the unit tests: the scripts' own parameter choices and grids (450 x 450, 48 x 2 x 48, 30 x 10 in a
shearing box with rotation, 128 x 128 with a passive tracer)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DEDALUS_REFERENCE", "/root/reference")
SAMPLES = [("samples/incompressible_hydro/swinging_wave/simulation.py", 6),
           ("samples/boussinesq_hydro/gravity_wave/2d_gmode_kx1_kz1.py", 4),
           ("samples/incompressible_hydro/2d_decaying_turbulence/2d_decaying_turbulence.py", 3),
           ("samples/incompressible_hydro/kelvin_helmholz/2d_kelvin_helmholz.py", 3),
           # Python-2 scripts: print statements converted in memory (oracle/build_ref.py's mechanical edits), nothing else
           ("samples/incompressible_mhd/alfven_wave/alfven_wave.py", 3),
           ("samples/incompressible_mhd/athena_field_loop/athena_field_loop_2d.py", 3)]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "samples")), reason="the reference tree is not present (GPU box)")
@pytest.mark.parametrize("script,iters", SAMPLES)
def test_reference_sample_script_runs_unmodified(tmp_path, script, iters):
    env = dict(os.environ, DDL_TEST_HOST_EMUL="1")
    env.pop("DEDALUS_DDL_LIB", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_sample.py"), os.path.join(REF, script), str(iters),
                        str(tmp_path)], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert r.returncode == 0 and "SAMPLE_OK iterations=%d" % iters in r.stdout, r.stdout[-3000:]
