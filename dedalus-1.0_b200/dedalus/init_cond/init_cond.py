"""Initial conditions (reference: dedalus/init_cond/init_cond.py).  Host-side, once per run:
written as tensor operations on the components' k-space buffers."""
import numpy as np
import torch

from ..analysis.volume_average import volume_average
from ..utils.logger import mylog
from ..utils.parallelism import com_sys


def _set(comp, index, value):
    comp.require_space("kspace")
    comp.kdata[tuple(index)] = value


def taylor_green(data):
    """Taylor-Green vortex.  The 2-D branch writes the very same eight entries as the reference
    (init_cond.py:43-51), including the four that land in the Nyquist-kx row of the transposed
    layout; the 3-D branch follows :52-85 (entries guarded by find_mode)."""
    mylog.info("Initializing Taylor Green Vortex.")
    u = data["u"]
    if u.ndim == 2:
        sx = {(1, 1): -1, (-1, -1): 1, (-1, 1): -1, (1, -1): 1}
        sy = {(1, 1): 1, (-1, -1): -1, (-1, 1): -1, (1, -1): 1}
        amp = 1j / 4.
    else:
        sx = {(1, 1, 1): -1, (1, 1, -1): 1, (-1, 1, 1): -1, (-1, 1, -1): 1,
              (1, -1, 1): -1, (1, -1, -1): 1, (-1, -1, 1): -1, (-1, -1, -1): 1}
        sy = {(1, 1, 1): 1, (1, 1, -1): 1, (-1, 1, 1): -1, (-1, 1, -1): -1,
              (1, -1, 1): 1, (1, -1, -1): 1, (-1, -1, 1): -1, (-1, -1, -1): -1}
        amp = 1j / 8.
    for comp, table in ((u["x"], sx), (u["y"], sy)):
        for idx, s in table.items():
            if u.ndim == 3:
                # the 3-D branch writes an entry only where find_mode() has the wavevector (:53-85): the
                # four kx = -1 entries do not exist in the half-complex layout and are skipped
                idx = comp.find_mode(idx)
                if idx is None:
                    continue
            _set(comp, idx, s * amp)


def sin_k(f, kindex, ampl=1.):
    f[tuple(kindex)] = ampl * 1j / 2.
    f[tuple(-1 * np.array(kindex))] = np.conj(ampl * 1j / 2.)


def cos_k(f, kindex, ampl=1.):
    f[tuple(kindex)] = ampl / 2.
    f[tuple(-1 * np.array(kindex))] = np.conj(ampl / 2.)


def alfven(data, k=(1, 0, 0), B0mag=5.0, u1mag=5e-6, p_vec=(0., 1., 0.)):
    """Alfven-wave initial condition (init_cond.py:146-232): uniform B0 along x plus a
    sinusoidal u, B perturbation at wavevector index k = (kz, ky, kx), polarisation p_vec."""
    if len(k) != 3:
        raise ValueError("Only setup for 3d")
    k = np.asarray(k, dtype=float)
    p_vec = np.asarray(p_vec, dtype=float)
    B0 = np.array([1., 0., 0.]) * B0mag
    c0 = data["u"]["x"]
    for i in range(3):
        data["B"][i].require_space("kspace")
        data["u"][i].require_space("kspace")
    zero = c0.find_mode((0., 0., 0.), exact=True)
    if zero is not None:
        for i in range(3):
            data["B"][i].kdata[zero] = B0[i]
    cA = B0mag / np.sqrt(4 * np.pi * data.parameters["rho0"])
    omega = np.abs(cA * np.dot(k, B0) / B0mag)
    u1 = p_vec * u1mag
    B1 = (np.dot(k, u1) * B0 - np.dot(k, B0) * u1) / omega
    for sign in (1.0, -1.0):
        # the reference compares k[2] with the z array, k[1] with y, k[0] with x (:177-183)
        idx = c0.find_mode((sign * k[1], sign * k[2], sign * k[0]), exact=True)
        if idx is None:
            continue
        for i in range(3):
            data["u"][i].kdata[idx] = sign * u1[i] * 1j / 2.
            data["B"][i].kdata[idx] = sign * B1[i] * 1j / 2.


def _uniform(shape, device, rng):
    """Uniform [0, 1) draws.  rng None: numpy's global generator on the host, exactly what the reference
    calls (`na.random.random`, init_cond.py:312-322,455), so `np.random.seed(s)` reproduces its fields
    number for number; rng "device": torch's generator on the GPU (no host pass; different numbers);
    else a numpy Generator / RandomState."""
    shape = tuple(int(n) for n in shape)
    if rng == "device":
        return torch.rand(shape, dtype=torch.float64, device=device)
    draw = np.random.random(shape) if rng is None else rng.random(shape)
    return torch.from_numpy(np.ascontiguousarray(draw)).to(device)


def turb_new(data, spec, tot_en=0.5, rng=None, **kwargs):
    """Random-phase solenoidal velocity field with spectrum `spec` (Rogallo 1981; reference
    init_cond.py:281-341).  `rng`: see _uniform (default: the reference's numpy global generator)."""
    c0 = data["u"][0]
    kk = torch.sqrt(c0.k2())
    kx, ky = c0.k["x"], c0.k["y"]
    sp = spec(kk, **kwargs)
    kk = torch.where(kk == 0, torch.ones_like(kk), kk)
    ampl = sp / (2. * np.pi * kk) if data.ndim == 2 else sp / (4 * np.pi * kk ** 2)
    aux = data.clone()
    aux.add_field("ampl", "ScalarField")
    aux["ampl"]["kspace"] = ampl
    aux["ampl"].dealias()
    norm = volume_average(aux["ampl"]["kspace"], kdict=c0.k, reduce_all=True)
    a = torch.sqrt(2. * aux["ampl"]["kspace"] * (tot_en / norm))
    eps = np.finfo(np.complex128).eps

    def random_phase():
        c0["xspace"] = _uniform(c0.local_shape["xspace"], c0._k.device, rng)
        th = c0["kspace"].clone()
        return th / torch.abs(th + eps)

    theta1, theta2, ph = random_phase(), random_phase(), random_phase()
    phi = torch.atan2(ph.imag, ph.real)
    alpha = a * theta1
    if data.ndim == 2:
        data["u"]["x"]["kspace"] = alpha * ky / kk
        data["u"]["y"]["kspace"] = -alpha * kx / kk
        if com_sys.myproc == 0:
            ux = data["u"]["x"].kdata
            ux[0, :] = (alpha * ky.abs() / kk)[0, :]
            ux[0, ux.shape[1] // 2 + 1] = 0.
    else:
        kz = c0.k["z"]
        kh = torch.sqrt(kx ** 2 + ky ** 2)
        kh = torch.where(kh == 0, torch.ones_like(kh), kh)
        alpha = alpha * torch.cos(phi)
        beta = a * theta2 * torch.sin(phi)
        data["u"]["x"]["kspace"] = (alpha * kk * ky + beta * kx * kz) / (kk * kh)
        data["u"]["y"]["kspace"] = (beta * ky * kz - alpha * kk * kx) / (kk * kh)
        data["u"]["z"]["kspace"] = -(beta * kh) / kk


def turb(ux, uy, spec, tot_en=0.5, rng=None, **kwargs):
    """The older 2-D random-phase generator on two components (init_cond.py:343-375): spectrum `spec`, one random phase
    array (numpy's global generator by default, like the reference), the kx = 0 column made Hermitian by hand.  `tot_en`
    is accepted and unused, as there."""
    kk = torch.zeros(ux._k.shape, dtype=torch.float64, device=ux._k.device)
    for kv in ux.k.values():
        kk = kk + kv ** 2
    kk = torch.sqrt(kk)
    k2 = torch.sqrt(ux.k["x"] ** 2 + ux.k["y"] ** 2)
    k2 = torch.where(k2 == 0, torch.ones_like(k2), k2)
    sp = spec(kk, **kwargs)
    kk = torch.where(kk == 0, torch.ones_like(kk), kk)
    ampl = torch.sqrt(sp / (2 * np.pi * kk))
    alpha = ampl * torch.exp(1j * 2 * np.pi * _uniform(ux._k.shape, ux._k.device, rng))
    a, b = ux.kdata, uy.kdata
    a[:, :] = alpha * kk * ux.k["y"] / (kk * k2)
    b[:, :] = -alpha * kk * ux.k["x"] / (kk * k2)
    n1 = a.shape[1]
    nh = n1 // 2 + 1
    if n1 % 2 == 0:
        start = nh - 2
        a[0, nh - 1] = 0.
    else:
        start = nh - 1
    idx = torch.arange(start, 0, -1, device=a.device)
    a[0, nh:] = a[0, idx].conj()
    b[0, nh:] = b[0, idx].conj()


def remove_compressible(ux, uy, renorm=False):
    """Project the compressive part off a 2-D velocity field (init_cond.py:377-389)."""
    if renorm:
        raise NotImplementedError
    ku = ux.kdata * ux.k["x"] + uy.kdata * uy.k["y"]
    ux.kdata.sub_(ku * ux.k["x"] / ux.k2(no_zero=True))
    uy.kdata.sub_(ku * uy.k["y"] / uy.k2(no_zero=True))


def MIT_vortices(data):
    """Three Gaussian vortices of the MIT 18.336 spectral NS demo (init_cond.py:391-408)."""
    y, x = data["u"]["x"].xspace_grid()
    aux = data.clone()
    aux.add_field("w", "ScalarField")
    aux.add_field("psi", "ScalarField")
    pi = np.pi
    aux["w"]["xspace"] = (torch.exp(-((x - pi) ** 2 + (y - pi + pi / 4) ** 2) / 0.2)
                          + torch.exp(-((x - pi) ** 2 + (y - pi - pi / 4) ** 2) / 0.2)
                          - 0.5 * torch.exp(-((x - pi - pi / 4) ** 2 + (y - pi - pi / 4) ** 2) / 0.4))
    aux["psi"]["kspace"] = aux["w"]["kspace"] / aux["w"].k2(no_zero=True)
    data["u"]["x"]["kspace"] = aux["psi"].deriv("y")
    data["u"]["y"]["kspace"] = -aux["psi"].deriv("x")


def vorticity_wave(data, mode, w_amp):
    """Single z-vorticity Fourier mode (init_cond.py:410-438)."""
    aux = data.clone()
    aux.add_field("w", "ScalarField")
    aux.add_field("psi", "ScalarField")
    i0 = data["u"]["x"].find_mode(mode)
    i1 = data["u"]["x"].find_mode(-1 * np.array(mode))
    if i0 is not None:
        aux["w"]["kspace"][i0] = w_amp / 2.
    if i1 is not None:
        aux["w"]["kspace"][i1] = np.conj(w_amp) / 2.
    if i0 is not None or i1 is not None:
        aux["psi"]["kspace"] = aux["w"]["kspace"] / aux["w"].k2(no_zero=True)
        data["u"]["x"]["kspace"] = aux["psi"].deriv("y")
        data["u"]["y"]["kspace"] = -aux["psi"].deriv("x")


def add_gaussian_white_noise(comp, std, rng=None):
    """Add a phasor of fixed amplitude and random phase to every mode, then restore Hermitian
    symmetry (init_cond.py:440-464).  `rng`: see _uniform."""
    phase = 2 * np.pi * _uniform(comp._k.shape, comp._k.device, rng)
    amp = std / np.sqrt(comp.nmodes - 1)
    noise = amp * torch.exp(1j * phase)
    zero = comp.find_mode([0.] * comp.ndim)
    if zero is not None:
        noise[zero] = 0.
    comp["kspace"].add_(noise)
    comp.enforce_hermitian()


def constant(data, comp, value):
    """Add a constant to a field ('T') or component ('ux') through its k = 0 mode (init_cond.py:466-486)."""
    if len(comp) == 1:
        rep = data[comp][0]
    elif len(comp) == 2:
        rep = data[comp[0]][comp[1]]
    else:
        raise ValueError("constant: comp %s invalid. must be 1 or 2 characters." % comp)
    zero = rep.find_mode([0.] * rep.ndim)
    if zero is None:
        return
    rep["kspace"][zero] += value + 0j
