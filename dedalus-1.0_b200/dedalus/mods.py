"""`from dedalus.mods import *` convenience namespace (reference: dedalus/mods.py:24-101)."""
from .config import decfg
from .utils.logger import mylog
from .analysis.api import (AnalysisSet, VolumeAverageSet, Snapshot, TrackMode, VolumeAverage, PowerSpectrum,
                           volume_average)
from .data_objects.api import FourierRepresentation, FourierShearRepresentation, ChebyshevRepresentation, StateData
from .init_cond.api import (taylor_green, sin_k, cos_k, turb, turb_new, mcwilliams_spec, MIT_vortices, vorticity_wave,
                            alfven, add_gaussian_white_noise, constant)
from .physics.api import IncompressibleHydro, BoussinesqHydro, IncompressibleMHD
from .time_stepping.api import RK2mid, RK2trap, RK4, CrankNicholsonVisc
from .utils.api import Timer, com_sys, swap_indices, restart
