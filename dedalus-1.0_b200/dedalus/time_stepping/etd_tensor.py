"""euler / etd1 / etd2rk1 / etd2rk2 as tensor operations with an integrating-factor ARRAY (reference:
dedalus/time_stepping/forward_step_cy_3d.pyx:17-124, forward_step_cy_2d.pyx:21-148).

The stage kernels of the hot path (include/ddl.h: ddl_stage) rebuild Z = -c (k^2)^n dt from the plan's static
wavenumber tables; in a shearing box k^2 depends on kx, ky AND time, so the compatibility path of
FourierShearRepresentation takes the factor as an array, exactly like the reference's Cython kernels, with their
branch structure: Z == 0 -> Euler form; |Z| < 0.5 -> the truncated series to Z^13 / 14!; else the closed forms
(3-D: f0 = exp(Z) in both branches; 2-D small-|Z| branch: f0 from the series, _2d:55)."""
import math

import torch

_FACT = [float(math.factorial(j)) for j in range(15)]


def _f_taylor(k, Z):
    """forward_step_cy_2d.pyx:123-148: sum_{j=k}^{14} Z^(j-k) / j!, ascending."""
    s = torch.zeros_like(Z)
    for j in range(k, 15):
        s = s + Z ** (j - k) / _FACT[j]
    return s


def _poly_desc(Z, kmin):
    """forward_step_cy_3d.pyx:55,87-88: explicit sums, powers by repeated multiplication, highest power first."""
    s = None
    for j in range(14, kmin - 1, -1):
        term = torch.ones_like(Z)
        for _ in range(j - kmin):
            term = term * Z
        term = term / _FACT[j]
        s = term if s is None else s + term
    return s


def _phis(Z, ndim):
    small = (Z < 0.5) & (Z > -0.5)
    Zs = torch.where(small, Z, torch.zeros_like(Z))
    Zl = torch.where(small, torch.ones_like(Z), Z)          # dummy 1 keeps the unused lanes finite
    e = torch.exp(Z)
    f1l = (torch.exp(Zl) - 1.0) / Zl
    f2l = (f1l - 1.0) / Zl
    if ndim == 2:
        return (torch.where(small, _f_taylor(0, Zs), e), torch.where(small, _f_taylor(1, Zs), f1l),
                torch.where(small, _f_taylor(2, Zs), f2l))
    return e, torch.where(small, _poly_desc(Zs, 1), f1l), torch.where(small, _poly_desc(Zs, 2), f2l)


def euler(start, output, deriv, dt):
    output.copy_(start + dt * deriv)


def etd1(start, output, deriv, intfactor, dt):
    Z = intfactor * dt
    f0, f1, _ = _phis(Z, start.dim())
    output.copy_(torch.where(Z == 0.0, start + dt * deriv, start * f0 + deriv * f1 * dt))


def etd2rk1(start, output, deriv1, deriv2, intfactor, dt):
    Z = intfactor * dt
    _, _, f2 = _phis(Z, start.dim())
    output.copy_(torch.where(Z == 0.0, start + dt / 2.0 * (deriv2 - deriv1), start + (deriv2 - deriv1) * f2 * dt))


def etd2rk2(start, output, deriv1, deriv2, intfactor, dt):
    Z = intfactor * dt
    f0, f1, f2 = _phis(Z, start.dim())
    output.copy_(torch.where(Z == 0.0, start + dt * deriv2, start * f0 + (deriv2 - deriv1) * 2.0 * f2 * dt + deriv1 * f1 * dt))
