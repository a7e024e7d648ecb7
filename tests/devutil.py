"""Helpers shared by the GPU tests: build device objects through the drop-in package and move
states between numpy (oracle / goldens) and the device."""
import ast
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return z, ast.literal_eval(str(z["meta"]))


def dev_physics(physics, shape, length=None, params=None, dealiasing="2/3 cython", direction=None):
    from dedalus.config import decfg
    import dedalus.physics.api as papi
    from dedalus.data_objects.api import FourierRepresentation
    decfg.set("FFT", "dealiasing", dealiasing)
    if direction is None:
        direction = "y" if len(shape) == 2 else "z"
    decfg.set("physics", "boussinesq_direction", direction)
    P = getattr(papi, physics)(tuple(shape), FourierRepresentation, tuple(length) if length else None)
    P.parameters.update(params or {})
    return P


def set_state(data, y):
    import torch
    for j, (_, _, c) in enumerate(data.components()):
        c["kspace"] = torch.from_numpy(np.ascontiguousarray(y[j]))


def get_state(data):
    return np.stack([c["kspace"].cpu().numpy() for _, _, c in data.components()])


def oracle_physics(physics, shape, length=None, params=None, dealiasing="2/3 cython", direction=None):
    import dedalus_oracle as orc
    kw = {}
    if physics == "BoussinesqHydro":
        kw["direction"] = direction or ("y" if len(shape) == 2 else "z")
    P = orc.PHYSICS[physics](shape, length, dealiasing, **kw)
    P.parameters.update(params or {})
    return P
