"""torchrun worker (gloo + host emulation, or NCCL on GPUs): the slab-decomposed package on the paths added late in round 1 --
the unfused helper sequence (FFT.dealiasing = None, '2/3 spherical') and a grid that is not a power of two -- against the oracle."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (os.path.join(ROOT, "dedalus-1.0_b200"), os.path.join(ROOT, "oracle"), HERE):
    sys.path.insert(0, p)
EMUL = os.environ.get("DDL_TEST_HOST_EMUL") == "1"


def main(out_path):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if EMUL:
        import conftest  # noqa: F401
        dist.init_process_group("gloo")
        dev = "cpu"
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dev = "cuda"
    import dedalus_oracle as orc
    from devutil import dev_physics, oracle_physics
    import dedalus.time_stepping.api as tapi
    from dedalus.config import decfg
    results = []
    for physics, shape, integ, dl in [("IncompressibleMHD", (16, 16, 16), "RK2mid", "None"),
                                      ("BoussinesqHydro", (16, 16, 16), "RK4", "2/3 spherical"),
                                      ("IncompressibleHydro", (12, 20, 24), "RK4", "2/3 cython")]:
        params = dict(nu=1e-2, eta=1e-2, kappa=1e-2)
        Po = oracle_physics(physics, shape, None, params, dealiasing=dl)
        do = orc.synthetic_ic(Po, 7)
        y0 = do.kvector()
        P = dev_physics(physics, shape, None, params, dealiasing=dl)
        data = P.create_fields(0.)
        comps = [c for _, _, c in data.components()]
        rows = comps[0].local_rows["kspace"]
        assert comps[0]._plan.nranks == world
        for j, c in enumerate(comps):
            c["kspace"] = torch.from_numpy(np.ascontiguousarray(y0[j][rows]))
        ti, to = getattr(tapi, integ)(P), orc.INTEGRATORS[integ](Po)
        for _ in range(2):
            ti.do_advance(data, 5e-3)
            to.do_advance(do, 5e-3)
        y1 = do.kvector()
        loc = np.stack([c["kspace"].cpu().numpy() for c in comps])
        num = torch.tensor([np.linalg.norm(loc - y1[:, rows]) ** 2, np.linalg.norm(y1[:, rows]) ** 2], dtype=torch.float64, device=dev)
        dist.all_reduce(num)
        dt_dev, dt_orc = P.compute_dt(data), Po.compute_dt(do)
        results.append({"physics": physics, "shape": shape, "dealiasing": dl, "rel": float(torch.sqrt(num[0] / num[1])),
                        "dt": float(dt_dev), "dt_oracle": float(dt_orc), "unfused": bool(P._unfused)})
        decfg.set("FFT", "dealiasing", "2/3 cython")
    # a hydro state with an entry outside the dealias mask: fused stage kernel + ddl_stage_outside, rank-local
    Po = oracle_physics("IncompressibleHydro", (16, 16, 32), None, dict(nu=0.05))
    do = orc.synthetic_ic(Po, 8)
    y0 = do.kvector()
    y0[:, 7, 3, 2] = 0.3 - 0.1j
    for j, (_, _, c) in enumerate(do.components()):
        c.kdata[...] = y0[j]
    P = dev_physics("IncompressibleHydro", (16, 16, 32), None, dict(nu=0.05))
    data = P.create_fields(0.)
    comps = [c for _, _, c in data.components()]
    rows = comps[0].local_rows["kspace"]
    for j, c in enumerate(comps):
        c["kspace"] = torch.from_numpy(np.ascontiguousarray(y0[j][rows]))
    ti, to = tapi.RK4(P), orc.RK4(Po)
    for _ in range(3):
        ti.do_advance(data, 5e-3)
        to.do_advance(do, 5e-3)
    y1 = do.kvector()
    loc = np.stack([c["kspace"].cpu().numpy() for c in comps])
    num = torch.tensor([np.linalg.norm(loc - y1[:, rows]) ** 2, np.linalg.norm(y1[:, rows]) ** 2], dtype=torch.float64, device=dev)
    dist.all_reduce(num)
    junk_here = torch.tensor([float(abs(loc[0][list(rows).index(7), 3, 2])) if 7 in list(rows) else 0.0], dtype=torch.float64, device=dev)
    dist.all_reduce(junk_here)
    results.append({"physics": "IncompressibleHydro", "shape": (16, 16, 32), "dealiasing": "junk", "rel": float(torch.sqrt(num[0] / num[1])),
                    "dt": 1.0, "dt_oracle": 1.0, "unfused": bool(P._unfused), "junk_left": float(junk_here[0]),
                    "fused_cached": len(getattr(ti, "_fuse_cache", {}))})
    # shearing box, slab-decomposed, against the oracle's restatement (pinned to the reference goldens by tests/test_oracle_shear.py)
    import dedalus.physics.api as papi
    from dedalus.data_objects.api import FourierShearRepresentation
    for physics, shape, integ, S, t0 in [("IncompressibleHydro", (8, 16, 16), "RK2mid", 1.5, 0.0),
                                         ("IncompressibleMHD", (16, 16, 16), "RK2trap", -1.0, 1.3),
                                         ("BoussinesqHydro", (8, 16, 24), "RK2mid", 1.0, 0.4)]:
        params = dict(nu=1e-2, eta=2e-2, kappa=1e-2)
        decfg.set("physics", "boussinesq_direction", "z")
        kw = {"direction": "z"} if physics == "BoussinesqHydro" else {}
        Po = orc.PHYSICS[physics](shape, None, "2/3 cython", shear=True, **kw)
        Po.parameters.update({k: v for k, v in params.items() if k in Po.parameters or k in ("nu",)})
        Po.parameters["shear_rate"] = S
        do = Po.create_fields(t0)
        ncomp = len(list(do.components()))
        noise = np.random.default_rng(41).standard_normal((ncomp,) + tuple(shape))
        j = 0
        for _, f in do:
            for _, c in f:
                c["xspace"] = noise[j]
                c["kspace"]
                j += 1
            if f.ncomp > 1:
                f.div_free()
        P = getattr(papi, physics)(tuple(shape), FourierShearRepresentation)
        P.parameters.update({k: v for k, v in params.items() if k in P.parameters})
        P.parameters["shear_rate"] = S
        data = P.create_fields(t0)
        comps = [c for _, _, c in data.components()]
        rows = comps[0].local_rows["kspace"]
        z0, nzl = int(comps[0].offset["xspace"]), int(comps[0].local_shape["xspace"][0])
        assert comps[0]._plan.nranks == world and comps[0]._plan.full_ky
        j = 0
        for _, f in data:
            for _, c in f:
                c["xspace"] = torch.from_numpy(np.ascontiguousarray(noise[j][z0:z0 + nzl]))
                c["kspace"]
                j += 1
            if f.ncomp > 1:
                f.div_free()
        y0 = do.kvector()
        loc = np.stack([c["kspace"].cpu().numpy() for c in comps])
        init = torch.tensor([np.linalg.norm(loc - y0[:, rows]) ** 2, np.linalg.norm(y0[:, rows]) ** 2], dtype=torch.float64, device=dev)
        dist.all_reduce(init)
        ti, to = getattr(tapi, integ)(P), orc.INTEGRATORS[integ](Po)
        for _ in range(2):
            ti.do_advance(data, 5e-3)
            to.do_advance(do, 5e-3)
        y1 = do.kvector()
        loc = np.stack([c["kspace"].cpu().numpy() for c in comps])
        num = torch.tensor([np.linalg.norm(loc - y1[:, rows]) ** 2, np.linalg.norm(y1[:, rows]) ** 2], dtype=torch.float64, device=dev)
        dist.all_reduce(num)
        dt_dev, dt_orc = P.compute_dt(data), Po.compute_dt(do)
        results.append({"physics": physics, "shape": shape, "dealiasing": "shear %g" % S, "rel": float(torch.sqrt(num[0] / num[1])),
                        "rel_init": float(torch.sqrt(init[0] / init[1])), "dt": float(dt_dev), "dt_oracle": float(dt_orc),
                        "unfused": bool(P._unfused)})
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(results, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
