"""Initial energy spectra (reference: dedalus/init_cond/turb_spectra.py)."""


def mcwilliams_spec(k, k0, E0):
    """McWilliams (1990, JFM 219:361) spectrum, unnormalised."""
    return k ** 6. / (k + 2. * k0) ** 18.
