from .init_cond import (taylor_green, sin_k, cos_k, turb, turb_new, remove_compressible, MIT_vortices, vorticity_wave, alfven,
                        add_gaussian_white_noise, constant)
from .turb_spectra import mcwilliams_spec
