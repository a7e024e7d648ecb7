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
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(results, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
