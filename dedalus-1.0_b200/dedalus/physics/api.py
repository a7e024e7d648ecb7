from .physics import Physics, IncompressibleHydro, BoussinesqHydro, IncompressibleMHD, IntegratingFactor
