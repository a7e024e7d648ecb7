"""Analysis tasks run from the main loop (reference: dedalus/analysis/analysis_set.py:41-839): AnalysisSet and the tasks
that turn the state into NUMBERS -- TrackMode (mode amplitudes to text files, :470-583), PowerSpectrum (shell-averaged 1-D
power spectra to text files, :585-816) and VolumeAverage (:818-839).  Snapshot (:79-468) draws matplotlib frames: plotting
is outside this backend's scope, so it is accepted with the reference's arguments and draws nothing (one warning), which keeps
the reference's sample scripts running unmodified.  The same holds for PowerSpectrum's `plot` switch.

The numbers come from device tensors: a tracked mode is one 16-byte read, a spectrum is binned on the device (one masked
sum per shell, off the hot path) and crosses to the host as `n` doubles."""
import numpy as np
import torch

from ..config import decfg
from ..utils.logger import mylog
from ..utils.parallelism import com_sys


class AnalysisSet(object):
    """:41-62."""

    def __init__(self, data, ti):
        self.data = data
        self.ti = ti
        self.tasks = []

    def add(self, task):
        self.tasks.append(task)
        task._an = self
        task._n = len(self.tasks)
        task.setup(self.data, self.ti.iteration)

    def run(self):
        for task in self.tasks:
            if self.ti.iteration % task.cadence == 0:
                task.run(self.data, self.ti.iteration)

    def cleanup(self):
        for task in self.tasks:
            task.cleanup(self.data, self.ti.iteration)


class AnalysisTask(object):
    """:64-76."""

    def __init__(self, cadence, *args, **kwargs):
        self.cadence = cadence

    def setup(self, data, iter):
        pass

    def run(self, data, iter):
        pass

    def cleanup(self, data, iter):
        pass


class Snapshot(AnalysisTask):
    """Image frames of field slices (:79-468).  Accepted, not drawn: no plotting in this backend."""

    _warned = False

    def __init__(self, cadence, space=None, axis=None, index=None, units=None, dpi=None, cmap=None, even_scale=False):
        self.cadence = cadence
        self.options = dict(space=space, axis=axis, index=index, units=units, dpi=dpi, cmap=cmap, even_scale=even_scale)

    def setup(self, data, it):
        if not Snapshot._warned and com_sys.myproc == 0:
            mylog.warning("Snapshot: image frames are not drawn by this backend (no plotting); use TimeStepBase.snapshot for data.")
            Snapshot._warned = True


def _sum_over_ranks(value):
    """What com_sys.comm.reduce(..., op=SUM, root=0) gives rank 0 (:556-559); serial runs have no communicator."""
    if com_sys.comm is None:
        return value
    import torch.distributed as dist
    t = torch.tensor([complex(value).real, complex(value).imag], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t)
    return complex(float(t[0]), float(t[1]))


class TrackMode(AnalysisTask):
    """Record the complex amplitude of given modes to <field><comp>_mode_amplitudes.dat (:470-583).  `modelist`: physical
    wavevectors in k-space axis order ((ky,kz,kx) / (kx,ky)); `indexlist`: local k-space indices (None on ranks without)."""

    def __init__(self, cadence, fieldlist=None, modelist=[], indexlist=[]):
        self.cadence = cadence
        self.fieldlist = fieldlist
        self.modelist = list(modelist)
        self.indexlist = list(indexlist)

    def _files(self, data):
        for fname in self.fieldlist:
            field = data[fname]
            for cindex, comp in field:
                yield fname + str(field.ctrans[cindex]), comp

    def setup(self, data, it):
        if self.fieldlist is None:
            self.fieldlist = list(data.fields.keys())
        columns = "# time"
        for mode in self.modelist:
            columns += "\t" + str([float(m) for m in mode])
        c0 = next(data.components())[2]
        for index in self.indexlist:
            column = "" if index is None else str((int(index[0]) + int(c0.offset["kspace"]),) + tuple(int(i) for i in index[1:]))
            columns += "\t" + column
        if com_sys.myproc == 0:
            for name, _ in self._files(data):
                with open("%s_mode_amplitudes.dat" % name, "a") as f:
                    f.write("# Dedalus Mode Amplitudes\n")
                    f.write(columns + "\n")

    def run(self, data, it):
        for name, comp in self._files(data):
            comp.require_space("kspace")
            amplitudes = []
            for mode in self.modelist:
                index = comp.find_mode(mode)
                amp = complex(comp._k[tuple(index)].item()) if index is not None else 0.
                amplitudes.append(_sum_over_ranks(amp))
            for index in self.indexlist:
                amp = 0. if index is None else complex(comp._k[tuple(index)].item())
                amplitudes.append(_sum_over_ranks(amp))
            if com_sys.myproc == 0:
                with open("%s_mode_amplitudes.dat" % name, "a") as f:
                    f.write("%s\t" % data.time + "\t".join(repr(a) for a in amplitudes) + "\n")


class PowerSpectrum(AnalysisTask):
    """Shell-averaged 1-D power spectra to <field>_power_spectra.dat (:585-816; binning :778-816)."""

    _warned = False

    def __init__(self, cadence, fieldlist=None, norm=1., write=True, plot=True, loglog=True, nyquistlines=False,
                 dealiasinglines=True, dpi=None):
        self.cadence = cadence
        self.fieldlist = fieldlist
        self.norm = norm
        self.write = write
        self.plot = plot
        self.dpi = dpi if dpi is not None else decfg.getint("analysis", "powerspectrum_dpi")

    def setup(self, data, it):
        if self.fieldlist is None:
            self.fieldlist = list(data.fields.keys())
        self.firstrun = [True] * len(self.fieldlist)
        if self.plot and not PowerSpectrum._warned and com_sys.myproc == 0:
            mylog.warning("PowerSpectrum: plots are not drawn by this backend; the spectra are written to text files.")
            PowerSpectrum._warned = True
        if self.write and com_sys.myproc == 0:
            for fname in self.fieldlist:
                with open("%s_power_spectra.dat" % fname, "w") as f:
                    f.write("# Dedalus Power Spectrum\n")

    def run(self, data, it):
        for col, fname in enumerate(self.fieldlist):
            k, spectrum = self._compute_spectrum(data[fname], self.norm)
            if self.write and com_sys.myproc == 0:
                with open("%s_power_spectra.dat" % fname, "a") as f:
                    if self.firstrun[col]:
                        f.write("time\t" + "\t".join(repr(float(ki)) for ki in k) + "\n")
                    f.write("%s\t" % data.time + "\t".join(repr(float(si)) for si in spectrum) + "\n")
            self.firstrun[col] = False

    def _compute_spectrum(self, field, norm):
        """(:778-816) power samples |f_i|^2 * 2 pi k (2-D) or 4 pi k^2 (3-D), averaged over the non-empty samples of each
        of n = min(N_min / 2, 100) equal shells of [0, k_max); returns bin centres and the averages."""
        c0 = field[0] if field.ncomp > 1 else field.components[0]
        kmag = torch.sqrt(c0.k2())
        power = torch.zeros_like(kmag)
        for _, comp in field:
            comp.require_space("kspace")
            power += comp._k.abs() ** 2
        if field.ndim == 2:
            power1d = norm * power * 2. * np.pi * kmag
        else:
            power1d = norm * power * 4. * np.pi * kmag ** 2
        kmax = np.sqrt(np.sum(np.asarray(c0.kny, dtype=float) ** 2))
        n = int(np.min([np.min(c0.global_shape["xspace"]) / 2., 100]))
        kbottom = np.linspace(0, kmax, n, endpoint=False)
        ktop = kbottom + kbottom[1]
        spectrum = torch.zeros(n, dtype=torch.float64, device=kmag.device)
        count = torch.zeros(n, dtype=torch.float64, device=kmag.device)
        nonzero = power1d != 0
        for i in range(n):
            mask = (kmag >= float(kbottom[i])) & (kmag < float(ktop[i])) & nonzero
            spectrum[i] = power1d[mask].sum()
            count[i] = mask.sum()
        if com_sys.nproc != 1:
            import torch.distributed as dist
            dist.all_reduce(spectrum)
            dist.all_reduce(count)
            if com_sys.myproc != 0:
                return (None, None)
        count[count == 0] = 1.
        return (kbottom + kbottom[1] / 2., (spectrum / count).cpu().numpy())


class VolumeAverage(AnalysisTask):
    """Run a VolumeAverageSet at a cadence (:818-839)."""

    def __init__(self, cadence, va):
        self.cadence = cadence
        self.volume_average_object = va

    def run(self, data, it):
        self.volume_average_object.run()
