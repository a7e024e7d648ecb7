"""Volume-averaged diagnostics (reference: dedalus/analysis/volume_average.py): the invariants
the parity gate compares (ekin, emag, divergence_sum, mag_div_sum, ...).  Reductions run on the
device; only the scalar result crosses to the host."""
import numpy as np
import torch

from ..utils.parallelism import com_sys, reduce_sum, reduce_mean, reduce_max


def volume_average(data, kdict=None, space="kspace", reduce_all=False):
    """Mean over the box.  In k-space the half-complex layout stores kx >= 0 only, so every
    plane but kx = 0 counts twice (volume_average.py:71-106)."""
    if space == "kspace":
        if hasattr(data, "k") and hasattr(data, "require_space"):
            k, values = data.k, data["kspace"]
        else:
            k, values = kdict, data
            if k is None:
                raise ValueError("volume_average: data is not a Dedalus representation, so you must pass k vectors via kdict")
        if values.dim() == 3:
            if float(k["x"].reshape(-1)[0]) == 0:
                local = 2 * values[..., 1:].sum() + values[..., 0].sum()
            else:
                local = 2 * values.sum()
        else:
            if float(k["x"].reshape(-1)[0]) == 0:
                local = 2 * values[1:, ...].sum() + values[0, ...].sum()
            else:
                local = 2 * values.sum()
        return reduce_sum(local, reduce_all=reduce_all)
    elif space == "xspace":
        values = data if torch.is_tensor(data) else data["xspace"]
        return reduce_mean(values)
    raise ValueError("volume_average: must be either xspace or kspace")


class VolumeAverageSet(object):
    """Time series of registered tasks, one text line per call to run() (volume_average.py:30-69)."""

    known_analysis = {}

    def __init__(self, data, filename="time_series.dat"):
        self.data = data
        self.filename = filename
        self.tasks = []
        self.scratch = None
        if com_sys.myproc == 0:
            self.outfile = open(self.filename, "a")
            self.outfile.write("# Dedalus Volume Average\n")
            self.outfile.write("# Column 0: time\n")

    def add(self, name, fmt, options={}):
        self.tasks.append((self.known_analysis[name], fmt, options))
        if com_sys.myproc == 0:
            self.outfile.write("# Column %i: %s\n" % (len(self.tasks), name))

    def run(self):
        line = ["%10.5f" % self.data.time]
        for f, fmt, kwargs in self.tasks:
            val = f(self.data, self.scratch, **kwargs)
            if com_sys.myproc == 0:
                line.append(fmt % val)
        if com_sys.myproc == 0:
            self.outfile.write("\t".join(line) + "\n")
            self.outfile.flush()

    @classmethod
    def register_task(cls, func):
        cls.known_analysis[func.__name__] = func
        return func


task = VolumeAverageSet.register_task


def _energy(field, space, reduce_all=False):
    if space == "kspace":
        acc = sum(0.5 * c["kspace"].abs() ** 2 for _, c in field)
        return volume_average(acc, kdict=field[0].k if field.ncomp > 1 else field.components[0].k, reduce_all=reduce_all)
    acc = sum(0.5 * c["xspace"] ** 2 for _, c in field)
    return volume_average(acc, space="xspace")


@task
def ekin(data, scratch=None, space="kspace", reduce_all=False):
    return _energy(data["u"], space, reduce_all)


@task
def emag(data, scratch=None, space="kspace", reduce_all=False):
    return _energy(data["B"], space, reduce_all)


def _mean_square(comp):
    k = comp["kspace"]
    return volume_average((k * k.conj()).real, kdict=comp.k)


@task
def ux2(data, scratch=None, space="kspace"):
    return _mean_square(data["u"]["x"])


@task
def uy2(data, scratch=None, space="kspace"):
    return _mean_square(data["u"]["y"])


@task
def uz2(data, scratch=None, space="kspace"):
    return _mean_square(data["u"]["z"])


@task
def bx2(data, scratch=None, space="kspace"):
    return _mean_square(data["B"]["x"])


@task
def by2(data, scratch=None, space="kspace"):
    return _mean_square(data["B"]["y"])


@task
def bz2(data, scratch=None, space="kspace"):
    return _mean_square(data["B"]["z"])


@task
def temp2(data, scratch=None, space="kspace"):
    return _mean_square(data["T"].components[0])


@task
def comp_mean(data, scratch, fname, cindex):
    return volume_average(data[fname][cindex]["xspace"], space="xspace")


@task
def enstrophy(data, scratch=None, space="kspace"):
    """2-D enstrophy 0.5 <w_z^2> (volume_average.py:190-199)."""
    w = data["u"]["y"].deriv("x") - data["u"]["x"].deriv("y")
    return volume_average(0.5 * w.abs() ** 2, kdict=data["u"]["x"].k)


@task
def energy_dissipation(data, scratch=None):
    u = data["u"]
    w2 = ((u["z"].deriv("y") - u["y"].deriv("z")).abs() ** 2 + (u["x"].deriv("z") - u["z"].deriv("x")).abs() ** 2
          + (u["y"].deriv("x") - u["x"].deriv("y")).abs() ** 2)
    return volume_average(2 * data.parameters["nu"] * 0.5 * w2, kdict=u["x"].k)


def _div(field):
    acc = 0
    for i, c in field:
        acc = acc + c.deriv(field.ctrans[i])
    return acc


@task
def divergence(data, scratch=None):
    return volume_average(_div(data["u"]), kdict=data["u"]["x"].k)


@task
def divergence_sum(data, scratch=None):
    """sum over stored modes of |i k . u| (volume_average.py:287-295)."""
    return _div(data["u"]).abs().sum().item()


@task
def mag_div(data, scratch=None):
    return volume_average(_div(data["B"]), kdict=data["B"]["x"].k)


@task
def mag_div_sum(data, scratch=None):
    return _div(data["B"]).abs().sum().item()


def _max_task(fname, cname):
    def f(data, scratch=None):
        return reduce_max(data[fname][cname]["xspace"])
    f.__name__ = ("u" if fname == "u" else "b") + cname + "_max"
    return task(f)


ux_max, uy_max, uz_max = _max_task("u", "x"), _max_task("u", "y"), _max_task("u", "z")
bx_max, by_max, bz_max = _max_task("B", "x"), _max_task("B", "y"), _max_task("B", "z")


@task
def vort_cenk(data, scratch=None):
    """Centroid wavenumber of McWilliams 1990 (volume_average.py:201-212)."""
    k2 = data["u"]["x"].k2(no_zero=True)
    en = sum(0.5 * c["kspace"].abs() ** 2 for _, c in data["u"])
    en[(0,) * en.dim()] = 0.
    return ((k2 ** 1.5 * en).sum() / (k2 * en).sum()).item()


def kinetic_helicity(data):
    """<u . curl u> (not in the reference; an extra invariant for the parity gate)."""
    u = data["u"]
    if u.ncomp != 3:
        return 0.0
    w = [u["z"].deriv("y") - u["y"].deriv("z"), u["x"].deriv("z") - u["z"].deriv("x"), u["y"].deriv("x") - u["x"].deriv("y")]
    acc = sum((u[i]["kspace"] * w[i].conj()).real for i in range(3))
    return volume_average(acc, kdict=u["x"].k)


def cross_helicity(data):
    """<u . B> (not in the reference)."""
    acc = sum((data["u"][i]["kspace"] * data["B"][i]["kspace"].conj()).real for i in range(data["u"].ncomp))
    return volume_average(acc, kdict=data["u"]["x"].k)
