"""torchrun worker (TEST INFRASTRUCTURE): ranks touch their buffers ASYMMETRICALLY before a step, the way callers do --
only one rank reads `.kdata`; init_cond.taylor_green (3-D) writes only on the ranks `find_mode` finds.  The knowledge bits
such accesses drop gate device collectives (invariants all_reduce, Hermitian all_gather), so every rank must take the same
decision (Physics.sync_knowledge) or the job hangs.  Runs over gloo with the host-emulation library, or nccl on GPUs."""
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
    import dedalus.init_cond.api as ic
    results = []

    def compare(comps, y1, rows):
        loc = np.stack([c["kspace"].cpu().numpy() for c in comps])
        t = torch.tensor([np.linalg.norm(loc - y1[:, rows]) ** 2, np.linalg.norm(y1[:, rows]) ** 2], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(torch.sqrt(t[0] / t[1]))

    # (1) MHD: every rank loads the same state; then ONLY the last rank looks at one of its buffers (and rank 0 at another)
    for integ in ("RK4", "RK2mid"):
        shape, params, dt = (16, 32, 32), dict(nu=1e-2, eta=1e-2), 2e-3
        Po = oracle_physics("IncompressibleMHD", shape, None, params)
        do = orc.synthetic_ic(Po, 21)
        y0 = do.kvector().copy()
        P = dev_physics("IncompressibleMHD", shape, None, params)
        data = P.create_fields(0.)
        comps = [c for _, _, c in data.components()]
        rows = comps[0].local_rows["kspace"]
        for j, c in enumerate(comps):
            c["kspace"] = torch.from_numpy(np.ascontiguousarray(y0[j][rows]))
        ti, to = getattr(tapi, integ)(P), getattr(orc, integ)(Po)
        ti.do_advance(data, dt)
        to.do_advance(do, dt)                       # after this step every rank knows everything about the state
        if rank == world - 1:
            comps[1].kdata                          # drops _clean / _sym / _soln of u_y on this rank only
        if rank == 0:
            float(comps[4]["kspace"].abs().max())   # "print a mode" on rank 0 only
        for _ in range(2):
            ti.do_advance(data, dt)
            to.do_advance(do, dt)
        results.append({"case": "one rank touches, " + integ, "rel": compare(comps, do.kvector(), rows)})

    # (2) the package's own 3-D Taylor-Green generator: writes on the ranks that own the eight modes only
    shape, params, dt = (16, 16, 16), dict(nu=0.05), 5e-3
    Po = oracle_physics("IncompressibleHydro", shape, None, params)
    do = Po.create_fields(0.)
    orc.taylor_green(do)
    P = dev_physics("IncompressibleHydro", shape, None, params)
    data = P.create_fields(0.)
    ic.taylor_green(data)
    comps = [c for _, _, c in data.components()]
    rows = comps[0].local_rows["kspace"]
    flags = sorted(set(repr(c._soln) for c in comps))
    ti, to = tapi.RK4(P), orc.RK4(Po)
    for _ in range(3):
        ti.do_advance(data, dt)
        to.do_advance(do, dt)
    results.append({"case": "taylor_green 3-D", "rel": compare(comps, do.kvector(), rows), "soln_flags_after_ic": flags})
    gathered = [None] * world
    dist.all_gather_object(gathered, flags)
    if rank == 0:
        results[-1]["soln_flags_by_rank"] = gathered
        with open(out_path, "w") as f:
            json.dump(results, f)
        print(json.dumps(results))
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
