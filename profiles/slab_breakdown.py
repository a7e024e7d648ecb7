#!/usr/bin/env python
"""Where a slab-decomposed RK4 step spends its time (run under torchrun, one rank per GPU):
   full step  |  same step with the all-to-all skipped (compute + launch overhead only; results are
   garbage, timing only)  |  the 60 exchanges of one step alone, back to back.
   python -m torch.distributed.run --nproc-per-node N profiles/slab_breakdown.py --grid 512"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dedalus-1.0_b200"))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--grid", dest="n", type=int, default=512)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bench
from dedalus.mods import RK4

P, data, dt = bench.make_state(a.n)
ti = RK4(P)
pipe = next(data.components())[2]._plan.pipeline


def timed(fn, reps):
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


def phases():
    """CUDA-event marks inside the four RHS evaluations of one step, summed per phase (ms)."""
    pipe.trace = []
    ti.do_advance(data, dt)
    torch.cuda.synchronize()
    tr, pipe.trace = pipe.trace, None
    out = {}
    for (l0, e0), (l1, e1) in zip(tr[:-1], tr[1:]):
        if l1 != "start":
            out[l1] = out.get(l1, 0.0) + e0.elapsed_time(e1)
    return {k: round(v, 3) for k, v in out.items()}


full = timed(lambda: ti.do_advance(data, dt), a.steps)
ph_full = phases()
p2p = pipe.exchange_kind in ("p2p", "peer")
b = None if p2p else pipe.buffers(6, 9)


def exchanges_only():
    for _ in range(4):
        if p2p:
            w = [pipe._exchange_p2p(f, True) for f in range(6)] + [pipe._exchange_p2p(f, False) for f in range(9)]
        else:
            w = [pipe._exchange(b["xs"][f], b["ks"][f], True) for f in range(6)]
            w += [pipe._exchange(b["ks"][f], b["xs"][f], False) for f in range(9)]
        for x in w:
            x.wait()


ex = timed(exchanges_only, a.steps)
pipe.skip_exchange = True
nocomm = timed(lambda: ti.do_advance(data, dt), a.steps)
ph_nocomm = phases()
pipe.skip_exchange = False
bytes_out = 4 * 15 * sum(pipe.to_peer[r] for r in range(world) if r != pipe.rank) * 16
if rank == 0:
    print(json.dumps({"n": a.n, "gpus": world, "exchange": pipe.exchange_kind, "step_ms": round(full, 3), "step_without_exchange_ms": round(nocomm, 3),
                      "exchanges_alone_ms": round(ex, 3), "bytes_out_per_gpu_per_step": bytes_out,
                      "exchange_gbs_per_direction": round(bytes_out / ex / 1e6, 1),
                      "rhs_phases_ms_per_step": ph_full, "rhs_phases_without_exchange": ph_nocomm}))
dist.destroy_process_group()
