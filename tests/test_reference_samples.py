"""CPU (build container only: needs /root/reference): the reference's OWN sample scripts executed against the drop-in package
(kernels through the host emulation) AND against the reference itself (oracle/_ref), same script file, same seed, a few
iterations -- and the two final spectral states compared.  The script files are never modified or copied (two Python-2
scripts get their print statements converted in memory).  What this exercises beyond the unit tests: the scripts' own parameter
choices, initial conditions and grids (450 x 450, 48 x 2 x 48, 30 x 10 in a shearing box with rotation, 128 x 128 with a passive
tracer, 32^3 and 128^2 MHD), `from dedalus.mods import *`, AnalysisSet with its tasks, the advance loop, finalize."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DEDALUS_REFERENCE", "/root/reference")
SAMPLES = [("samples/incompressible_hydro/swinging_wave/simulation.py", 6),
           ("samples/boussinesq_hydro/gravity_wave/2d_gmode_kx1_kz1.py", 4),
           ("samples/incompressible_hydro/2d_decaying_turbulence/2d_decaying_turbulence.py", 3),
           ("samples/incompressible_hydro/kelvin_helmholz/2d_kelvin_helmholz.py", 3),
           ("samples/incompressible_mhd/alfven_wave/alfven_wave.py", 3),
           ("samples/incompressible_mhd/athena_field_loop/athena_field_loop_2d.py", 3)]


def run(script, iters, workdir, impl):
    os.makedirs(workdir, exist_ok=True)
    out = os.path.join(workdir, "final.npz")
    env = dict(os.environ, DDL_TEST_HOST_EMUL="1", OMP_NUM_THREADS="2")
    env.pop("DEDALUS_DDL_LIB", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_sample.py"), os.path.join(REF, script), str(iters),
                        workdir, impl, out], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert r.returncode == 0 and "SAMPLE_OK iterations=%d" % iters in r.stdout, (impl, r.stdout[-3000:])
    return np.load(out)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "samples")), reason="the reference tree is not present (GPU box)")
@pytest.mark.parametrize("script,iters", SAMPLES)
def test_sample_script_same_result_as_the_reference(tmp_path, script, iters):
    ours = run(script, iters, str(tmp_path / "ours"), "ours")
    ref = run(script, iters, str(tmp_path / "ref"), "ref")
    assert int(ours["iteration"]) == int(ref["iteration"]) == iters
    assert abs(float(ours["time"]) - float(ref["time"])) <= 1e-13 * max(1.0, abs(float(ref["time"])))
    err = np.linalg.norm(ours["state"] - ref["state"]) / np.linalg.norm(ref["state"])
    assert err < 1e-10, err
    # the scripts' states are dominated by a background (B0, a single wave): compare what the steps CHANGED as well
    assert np.linalg.norm(ours["state0"] - ref["state0"]) <= 1e-13 * np.linalg.norm(ref["state0"])
    d_ours, d_ref = ours["state"] - ours["state0"], ref["state"] - ref["state0"]
    assert np.linalg.norm(d_ref) > 0
    assert np.linalg.norm(d_ours - d_ref) < 1e-9 * np.linalg.norm(d_ref), np.linalg.norm(d_ours - d_ref) / np.linalg.norm(d_ref)
