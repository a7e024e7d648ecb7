"""Worker of tests/test_slab_gloo.py: one process = one rank of a slab decomposition, driving
the product's SlabPipeline (dedalus/data_objects/slab.py) with the HOST-EMULATION library and
CPU tensors over gloo.  Checks the host-side logic of the N>1 path (partition, peer-blocked
layouts, split sizes, pipelined exchange order) against the reference goldens."""
import ast
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (os.path.join(HERE, "host"), os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def _load_slab_module():
    """dedalus/data_objects/slab.py on its own (the package __init__ would load the CUDA library)."""
    path = os.path.join(ROOT, "dedalus-1.0_b200", "dedalus", "data_objects", "slab.py")
    spec = importlib.util.spec_from_file_location("ddl_slab_under_test", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def worker(rank, world, port, case, out_path, layout=0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import dedalus_oracle as orc
        import emul
        slab = _load_slab_module()
        lib = emul.load()
        sys.path.insert(0, os.path.join(ROOT, "dedalus-1.0_b200", "dedalus", "_lib"))
        z = np.load(os.path.join(HERE, "golden", case + ".npz"))
        meta = ast.literal_eval(str(z["meta"]))
        g = orc.Grid(meta["shape"], meta["length"], meta.get("dealiasing", "2/3 cython"))
        pl = emul.EmulPlan(lib, g, world, rank, layout)
        # argtypes as the product binding sets them
        import ctypes as C
        vp, i32 = C.c_void_p, C.c_int
        lib.ddl_slab_info.argtypes = [vp, vp]; lib.ddl_slab_rows.argtypes = [vp, vp]
        for name in ("ddl_slab_zinv", "ddl_slab_yinv", "ddl_slab_yfwd"):
            getattr(lib, name).argtypes = [vp, i32, vp, vp, vp]
        lib.ddl_slab_xfused.argtypes = [vp, i32, vp, vp, vp, vp]
        lib.ddl_slab_xc2r.argtypes = [vp, vp, vp, vp]; lib.ddl_slab_xr2c.argtypes = [vp, vp, vp, vp]
        lib.ddl_slab_zfwd.argtypes = [vp, i32, vp, vp, i32, vp]
        lib.ddl_slab_assemble.argtypes = [vp, i32, vp, vp, vp, vp, vp]
        lib.ddl_dealias.argtypes = [vp, vp, vp]
        pipe = slab.SlabPipeline(lib, pl.plan, torch.device("cpu"))
        nz, ny, nx = g.shape
        nyl, nzl = ny // world, nz // world
        assert (pipe.P, pipe.rank, pipe.nzl, pipe.nyl) == (world, rank, nzl, nyl)
        ksl = slice(rank * nyl, (rank + 1) * nyl) if layout == 0 else slice(rank, ny, world)
        zsl = slice(rank * nzl, (rank + 1) * nzl)
        res = {}
        # ---- transforms: backward / forward of one component against the oracle
        y0 = z["y0"]
        c = orc.Comp(g)
        c["kspace"] = y0[0]
        xref = c["xspace"].copy()
        kdeal = c.kdata.copy()
        k_loc = torch.from_numpy(np.ascontiguousarray(y0[0][ksl]))
        x_loc = torch.full((nzl, ny, nx), float("nan"), dtype=torch.float64)
        pipe.backward(k_loc, x_loc)
        res["bwd"] = rel(x_loc.numpy(), xref[zsl])
        res["bwd_dealias"] = rel(k_loc.numpy(), kdeal[ksl])
        k2 = torch.full((nyl, nz, nx // 2 + 1), complex("nan"), dtype=torch.complex128)
        k2.zero_()
        pipe.forward(x_loc, k2)
        c["kspace"]
        res["fwd"] = rel(k2.numpy(), c.kdata[ksl])
        # ---- fused RHS against the reference golden
        pid = emul.EmulPlan.PHYS[meta["physics"]]
        prm = dict(meta["params"])
        pp = emul.PhysParams(prm.get("rho0", 1.0), prm.get("g", 1.0), prm.get("alpha_t", 1.0), prm.get("beta", 1.0), 2, 0)
        state = [torch.from_numpy(np.ascontiguousarray(s[ksl])) for s in y0]
        deriv = [torch.full_like(s, complex("nan")) for s in state]
        for d in deriv:
            d.zero_()
        pipe.rhs(pid, pp, state, deriv, meta["physics"] == "IncompressibleMHD", False)
        d = np.stack([t.numpy() for t in deriv])
        res["rhs"] = rel(d, z["dy0"][:, ksl])
        res["state_after"] = rel(np.stack([t.numpy() for t in state]), z["y0_after_rhs"][:, ksl])
        res["exchanges"] = pipe.exchanges
        res["rows"] = pipe.rows
        gathered = [None] * world
        dist.all_gather_object(gathered, res)
        if rank == 0:
            import json
            with open(out_path, "w") as f:
                json.dump(gathered, f)
    finally:
        dist.destroy_process_group()
