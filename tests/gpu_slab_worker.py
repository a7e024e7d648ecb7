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


EMUL = os.environ.get("DDL_TEST_HOST_EMUL") == "1"    # tests/conftest.py host-emulation harness: gloo + CPU tensors


def main(out_path):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if EMUL:
        import conftest  # noqa: F401  (points the package at the host-emulation library, in this process only)
        dist.init_process_group("gloo")
        dev = "cpu"
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dev = "cuda"
    import dedalus_oracle as orc
    from devutil import rel, dev_physics, oracle_physics
    import dedalus.time_stepping.api as tapi
    import dedalus.analysis.volume_average as va
    results = []
    def make_ic(Po, compressive):
        if not compressive:
            return orc.synthetic_ic(Po, 11)
        # NOT solenoidal: the package must notice and take the advective-form policies (DDL_*_ADV), which use
        # the collective exchange whatever DEDALUS_SLAB_EXCHANGE says
        do = Po.create_fields(0.)
        rng = np.random.default_rng(29)
        for _, _, c in do.components():
            c["xspace"] = 0.3 * rng.standard_normal(Po.g.shape)
            c.require_space("kspace")
        return do

    for physics, shape, integ, nsteps, params, compressive in [
            ("IncompressibleMHD", (32, 32, 64), "RK4", 3, dict(nu=1e-3, eta=2e-3), False),
            ("BoussinesqHydro", (16, 32, 32), "RK2mid", 3, dict(nu=1e-3, kappa=1e-3), False),
            ("IncompressibleHydro", (32, 16, 128), "RK4", 2, dict(nu=1e-3), False),
            ("IncompressibleMHD", (16, 32, 32), "RK4", 2, dict(nu=1e-2, eta=1e-2), True),
            ("BoussinesqHydro", (16, 16, 32), "RK2trap", 2, dict(nu=1e-2, kappa=1e-2), True)]:
        Po = oracle_physics(physics, shape, None, params)
        do = make_ic(Po, compressive)
        y0 = do.kvector()
        ko = Po.create_fields(0.)
        Po.RHS(do, ko)
        dy0 = ko.kvector()
        do = Po.create_fields(0.)
        for (_, _, c), y in zip(do.components(), y0):
            c["kspace"] = y
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
                           dtype=torch.float64, device=dev)
        mx = num[2:].clone()
        dist.all_reduce(num, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        ek = va.ekin(data, reduce_all=True)
        ek_orc = orc.energy(do, "u")
        # CFL limit from the reduction inside the x pass (slab: per-rank maxima, all-reduced),
        # then one advance() that takes its dt from its own first RHS evaluation
        dt_dev, dt_orc = P.compute_dt(data), Po.compute_dt(do)
        ti.CFL, ti.iteration, ti.save_cadence, ti.max_save_period = 0.3, 1, 10 ** 9, 1e300
        ti.advance(data)
        verdict = bool(all(c._soln for n, _, c in data.components() if n in ("u", "B")))     # before any buffer is handed out
        to.do_advance(do, 0.3 * dt_orc)
        y2 = do.kvector()
        loc = np.stack([c["kspace"].cpu().numpy() for c in comps])
        num2 = torch.tensor([np.linalg.norm(loc - y2[:, rows]) ** 2, np.linalg.norm(y2[:, rows]) ** 2], dtype=torch.float64, device=dev)
        dist.all_reduce(num2, op=dist.ReduceOp.SUM)
        emag = va.emag(data, reduce_all=True) if physics == "IncompressibleMHD" else 0.0
        results.append({"physics": physics, "shape": shape, "world": world, "compressive": compressive,
                        "solenoidal_verdict": verdict, "rel_vs_oracle": float(torch.sqrt(num[0] / num[1])),
                        "rhs_rel": float(mx[0]), "ekin": float(ek), "ekin_oracle": float(ek_orc),
                        "dt": float(dt_dev), "dt_oracle": float(dt_orc), "rel_after_cfl_step": float(torch.sqrt(num2[0] / num2[1])),
                        "dt_taken": float(ti.dt_old), "emag": float(emag),
                        "emag_oracle": float(orc.energy(do, "B")) if physics == "IncompressibleMHD" else 0.0,
                        "exchanges": comps[0]._plan.pipeline.exchanges, "ky_layout": comps[0]._plan.ky_layout,
                        "last_rhs_path": getattr(comps[0]._plan.pipeline, "last_rhs_path", None),
                        "exchange": comps[0]._plan.pipeline.exchange_kind})
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(results, f)
        print(json.dumps(results))
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
