"""Child process of tests/test_reference_samples.py: execute one of the REFERENCE's own sample scripts against
  --impl ours : the drop-in package (the script's `from dedalus.mods import *` resolves to dedalus-1.0_b200/dedalus; kernels
                through the host emulation when DDL_TEST_HOST_EMUL=1), or
  --impl ref  : the reference itself (oracle/_ref: its physics, representations, initial conditions, integrators and
                volume averages; its plotting tasks, which need matplotlib, and its HDF5 snapshots, which need h5py, are
                replaced by no-ops, and the shearing box's inverse route by the restatement of rev_fftw that
                tests/golden/make_shear_goldens.py documents),
stopping after a few iterations, and save the final spectral state.  The script file itself is never modified; a Python-2 script
gets the mechanical edits of oracle/build_ref.py (print statements, xrange, ...) in memory.
usage: run_reference_sample.py <script> <max_iterations> <workdir> <impl> <out.npz>"""
import os
import sys
import types

script, max_iter, workdir, impl, out = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402

if impl == "ours":
    sys.path.insert(0, HERE)
    import conftest  # noqa: F401   (paths; the host-emulation harness when DDL_TEST_HOST_EMUL=1)
    import dedalus.time_stepping.time_step as ts

    def to_numpy(a):
        return a.cpu().numpy()
else:
    import build_ref
    decfg, data_api, physics_api, ts = build_ref.import_ref()
    import numpy.fft as npfft
    from dedalus.data_objects.api import FourierRepresentation, FourierShearRepresentation
    import dedalus.init_cond.init_cond as ic
    from dedalus.init_cond.turb_spectra import mcwilliams_spec
    import dedalus.analysis.volume_average as va
    from dedalus.utils.parallelism import com_sys, swap_indices
    from dedalus.utils.timer import Timer
    from dedalus.utils.logger import mylog

    def rev_like_fftw(self):          # see tests/golden/make_shear_goldens.py
        shape = self.global_shape['xspace']
        if self.ndim == 2:
            k = npfft.ifft(self.kdata, axis=1) * shape[0]
        else:
            k = npfft.ifftn(self.kdata, axes=(0, 1)) * (shape[0] * shape[1])
        self._mdata[:] = np.transpose(k, [1, 0, 2][:self.ndim])
        self._mdata *= np.exp(-1j * self._phase_rate * self.sd.time)
        self.xdata[:] = npfft.irfft(self._mdata, n=int(shape[-1]), axis=-1) * shape[-1]

    FourierShearRepresentation.rev_np = rev_like_fftw
    ts.TimeStepBase.snapshot = lambda self, data: None

    class _Task(object):
        def __init__(self, cadence=1, *a, **kw):
            self.cadence = cadence

    class AnalysisSet(object):
        def __init__(self, data, ti):
            self.tasks, self.ti = [], ti

        def add(self, task):
            self.tasks.append(task)

        def run(self):
            for t in self.tasks:
                if isinstance(t, VolumeAverage) and self.ti.iteration % t.cadence == 0:
                    t.va.run()

        def cleanup(self):
            pass

    class VolumeAverage(_Task):
        def __init__(self, cadence, va_obj):
            self.cadence, self.va = cadence, va_obj

    mods = types.ModuleType("dedalus.mods")
    for name in ("IncompressibleHydro", "BoussinesqHydro", "IncompressibleMHD"):
        setattr(mods, name, getattr(physics_api, name))
    for name in ("RK2mid", "RK2trap", "RK4", "CrankNicholsonVisc"):
        setattr(mods, name, getattr(ts, name))
    for name in ("taylor_green", "sin_k", "cos_k", "turb_new", "MIT_vortices", "vorticity_wave", "alfven", "add_gaussian_white_noise",
                 "constant"):
        setattr(mods, name, getattr(ic, name))
    mods.__dict__.update(decfg=decfg, mylog=mylog, FourierRepresentation=FourierRepresentation,
                         FourierShearRepresentation=FourierShearRepresentation, mcwilliams_spec=mcwilliams_spec,
                         VolumeAverageSet=va.VolumeAverageSet, com_sys=com_sys, swap_indices=swap_indices, Timer=Timer,
                         AnalysisSet=AnalysisSet, VolumeAverage=VolumeAverage, Snapshot=_Task, TrackMode=_Task, PowerSpectrum=_Task)
    sys.modules["dedalus.mods"] = mods

    def to_numpy(a):
        return np.array(a)

_ok = ts.TimeStepBase.ok
ts.TimeStepBase.ok = property(lambda self: self.iteration < max_iter and _ok.fget(self))
_advance = ts.TimeStepBase.advance
_initial = []


def stack(data):
    return np.stack([to_numpy(c["kspace"]).copy() for fn, f in data for i, c in f])


def advance(self, data, *a, **kw):
    if not _initial:
        _initial.append(stack(data))           # the state the script's own setup produced, before the first step
    return _advance(self, data, *a, **kw)


ts.TimeStepBase.advance = advance
os.chdir(workdir)
np.random.seed(20261017)           # turb_new / white noise draw from numpy's global generator in both implementations
src = open(script).read()
try:
    code = compile(src, script, "exec")
except SyntaxError:
    import build_ref
    code = compile(build_ref.transliterate("sample", src), script, "exec")
ns = {"__name__": "__main__", "__file__": script}
exec(code, ns)
data, ti = ns["data"], ns["ti"]
state = stack(data)
assert np.isfinite(state).all() and np.linalg.norm(state) > 0
np.savez(out, state=state, state0=_initial[0], time=ti.time, iteration=ti.iteration)
print("SAMPLE_OK iterations=%d time=%.6e norm=%.12e" % (ti.iteration, ti.time, np.linalg.norm(state)))
