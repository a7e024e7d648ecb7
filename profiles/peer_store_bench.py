#!/usr/bin/env python
"""NVLink throughput of the peer-store passes in isolation (torchrun, one rank per GPU):
   inverse z pass storing into the peers' arenas, per field and as one 6-field launch, and the same
   launches with every block redirected to LOCAL memory (what the pass costs without NVLink)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dedalus-1.0_b200"))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
os.environ.setdefault("DEDALUS_KY_LAYOUT", "cyclic")
import bench
from dedalus.mods import RK4
from dedalus.data_objects.slab import _ptrs

P, data, dt = bench.make_state(512)
ti = RK4(P)
ti.do_advance(data, dt)
pipe = next(data.components())[2]._plan.pipeline
lib, h = pipe.lib, pipe.h
state = [c._k for _, _, c in data.components()]
b = pipe._p2p_buffers(6, 9)
st = torch.cuda.current_stream().cuda_stream
zt, yt = pipe._zinv_tab, pipe._yfwd_tab
# local redirection: every block of field f -> this rank's own x-side / k-side field f
zloc = zt[:, rank:rank + 1].expand(-1, world).contiguous()
yloc = yt[:, rank:rank + 1].expand(-1, world).contiguous()
el = 8 * world


def timed(fn, reps=5):
    fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def zinv_per_field(tab):
    for f in range(6):
        pipe._check(lib.ddl_slab_zinv_peer(h, 1, _ptrs([state[f]]), tab.data_ptr() + f * el, st))


def zinv_all(tab):
    pipe._check(lib.ddl_slab_zinv_peer(h, 6, _ptrs(state), tab.data_ptr(), st))


def yfwd_all(tab):
    pipe._check(lib.ddl_slab_yfwd_peer(h, 9, _ptrs(b["c"][:9]), tab.data_ptr(), 0, pipe.nzl, st))


blk = pipe.nzl * pipe.cx * 16
out_inv = 6 * sum(pipe.cyl * blk for r in range(world) if r != rank)
out_fwd = 9 * sum(pipe.rows[r] * blk for r in range(world) if r != rank)
res = {"gpus": world}
for name, fn, tab, nb in (("zinv_per_field_peer", zinv_per_field, zt, out_inv), ("zinv_one_launch_peer", zinv_all, zt, out_inv),
                          ("zinv_one_launch_local", zinv_all, zloc, 0), ("yfwd_one_launch_peer", yfwd_all, yt, out_fwd),
                          ("yfwd_one_launch_local", yfwd_all, yloc, 0)):
    ms = timed(lambda: fn(tab))
    res[name] = {"ms": round(ms, 3), "nvlink_gbs_out": round(nb / ms / 1e6, 1) if nb else None}
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
