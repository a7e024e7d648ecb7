"""torchrun worker of tests/test_gpu_slab.py: every rank advances its slab of the same global
state through the drop-in API; rank 0 compares the gathered result with the oracle."""
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


def main(out_path):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import dedalus_oracle as orc
    from devutil import rel, dev_physics, oracle_physics
    import dedalus.time_stepping.api as tapi
    import dedalus.analysis.volume_average as va
    results = []
    for physics, shape, integ, nsteps, params in [
            ("IncompressibleMHD", (32, 32, 64), "RK4", 3, dict(nu=1e-3, eta=2e-3)),
            ("BoussinesqHydro", (16, 32, 32), "RK2mid", 3, dict(nu=1e-3, kappa=1e-3)),
            ("IncompressibleHydro", (32, 16, 128), "RK4", 2, dict(nu=1e-3))]:
        Po = oracle_physics(physics, shape, None, params)
        do = orc.synthetic_ic(Po, 11)
        y0 = do.kvector()
        ko = Po.create_fields(0.)
        Po.RHS(do, ko)
        dy0 = ko.kvector()
        do = orc.synthetic_ic(Po, 11)
        P = dev_physics(physics, shape, None, params)
        data, deriv = P.create_fields(0.), P.create_fields(0.)
        comps = [c for _, _, c in data.components()]
        rows, nyl = comps[0].local_rows["kspace"], int(comps[0].local_shape["kspace"][0])
        assert nyl == shape[1] // world and comps[0]._plan.nranks == world
        for j, c in enumerate(comps):
            c["kspace"] = torch.from_numpy(np.ascontiguousarray(y0[j][rows]))
        P.RHS(data, deriv)
        d_loc = np.stack([c["kspace"].cpu().numpy() for _, _, c in deriv.components()])
        rhs_rel = rel(d_loc, dy0[:, rows])
        for j, c in enumerate(comps):
            c["kspace"] = torch.from_numpy(np.ascontiguousarray(y0[j][rows]))
        dt = 2e-3
        ti = getattr(tapi, integ)(P)
        to = getattr(orc, integ)(Po)
        for _ in range(nsteps):
            ti.do_advance(data, dt)
            to.do_advance(do, dt)
        y1 = do.kvector()
        loc = np.stack([c["kspace"].cpu().numpy() for c in comps])
        num = torch.tensor([np.linalg.norm(loc - y1[:, rows]) ** 2, np.linalg.norm(y1[:, rows]) ** 2, rhs_rel],
                           dtype=torch.float64, device="cuda")
        mx = num[2:].clone()
        dist.all_reduce(num, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        ek = va.ekin(data, reduce_all=True)
        results.append({"physics": physics, "shape": shape, "world": world, "rel_vs_oracle": float(torch.sqrt(num[0] / num[1])),
                        "rhs_rel": float(mx[0]), "ekin": float(ek), "ekin_oracle": float(orc.energy(do, "u")),
                        "exchanges": comps[0]._plan.pipeline.exchanges, "ky_layout": comps[0]._plan.ky_layout,
                        "exchange": comps[0]._plan.pipeline.exchange_kind})
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(results, f)
        print(json.dumps(results))
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
