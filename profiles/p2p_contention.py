#!/usr/bin/env python
"""How fast do the peer pushes of the slab exchange go while the GPU is busy with HBM-bound kernels?
   torchrun --nproc-per-node N profiles/p2p_contention.py
Times 15 exchanges (one RHS worth) alone and concurrently with a stream of large device copies."""
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
import bench
from dedalus.mods import RK4

P, data, dt = bench.make_state(512)
ti = RK4(P)
ti.do_advance(data, dt)
pipe = next(data.components())[2]._plan.pipeline
side = torch.cuda.Stream()
a = torch.empty(1 << 28, dtype=torch.float64, device="cuda")   # 2 GiB
b = torch.empty_like(a)


def exchanges():
    w = [pipe._exchange_p2p(f, True) for f in range(6)] + [pipe._exchange_p2p(f, False) for f in range(9)]
    for x in w:
        x.wait()


def run(load):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if load:
        l0.record()
        for _ in range(12):
            b.copy_(a)
        l1.record()
    with torch.cuda.stream(side):
        e0.record()
        for _ in range(4):
            exchanges()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), (l0.elapsed_time(l1) if load else 0.0)


run(False)
alone, _ = run(False)
busy, load_ms = run(True)
t = torch.tensor([alone, busy, load_ms], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    nbytes = 4 * 15 * sum(pipe.to_peer[r] for r in range(world) if r != pipe.rank) * 16
    print(json.dumps({"gpus": world, "exchanges_alone_ms": round(float(t[0]), 3), "exchanges_under_load_ms": round(float(t[1]), 3),
                      "load_ms": round(float(t[2]), 3), "load_gbs": round(12 * 2 * a.numel() * 8 / float(t[2]) / 1e6, 1),
                      "gbs_alone": round(nbytes / float(t[0]) / 1e6, 1), "gbs_under_load": round(nbytes / float(t[1]) / 1e6, 1)}))
dist.destroy_process_group()
