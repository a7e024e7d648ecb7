#!/usr/bin/env python
"""Multi-GPU schedule sweep of the headline step (3-D MHD RK4, slab-decomposed, peer exchange) in ONE torchrun launch:
   torchrun --nproc-per-node 8 profiles/slab_sweep.py [--n 512] [--steps 6]
The state is built once; every configuration of the slab pipeline's knobs (plane chunks of the forward half, field groups of
the inverse half, CTA limit of the NVLink-bound peer-store passes, priority of the stream they run on) is warmed up and timed
with CUDA events (max over ranks).  Every configuration must leave the same energies behind (same arithmetic, other schedule)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dedalus-1.0_b200"))

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--configs", default="")
a = ap.parse_args()

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
os.environ.setdefault("DEDALUS_KY_LAYOUT", "cyclic")
os.environ.setdefault("DEDALUS_SLAB_EXCHANGE", "peer")
import bench
import dedalus._lib as L
from dedalus.mods import RK4
import dedalus.analysis.volume_average as va

P, data, dt = bench.make_state(a.n)
ti = RK4(P)
for _ in range(3):
    ti.do_advance(data, dt)
pipe = next(data.components())[2]._plan.pipeline
nk = a.n * a.n * (a.n // 2 + 1)

# (chunks, inverse, peer_ctas [peer] / push_ctas [push], side_priority[, exchange kind = "peer"])
CONFIGS = [(4, "batched", 0, 0), (2, "batched", 0, 0), (4, "batched", 64, 1), (4, "batched", 74, 1), (4, "batched", 96, 1), (4, "batched", 110, 1),
           (8, "batched", 74, 1), (8, "batched", 96, 1), (8, "groups:2", 74, 1), (8, "groups:2", 96, 1), (8, "groups:3", 74, 1), (8, "groups:3", 96, 1),
           (16, "groups:2", 96, 1), (4, "groups:2", 96, 1), (8, "groups:2", 96, 0), (8, "groups:2", 120, 1),
           (4, "groups:3", 96, 1, "push"), (4, "groups:3", 128, 1, "push"), (8, "groups:3", 148, 1, "push")]
_OLD = [(4, "batched", 0, 0), (2, "batched", 0, 0), (8, "batched", 0, 0), (16, "batched", 0, 0), (4, "groups:2", 0, 0),
           (4, "groups:3", 0, 0), (4, "fields", 0, 0), (8, "groups:2", 0, 0), (4, "batched", 0, 1), (8, "groups:2", 0, 1),
           (4, "batched", 148, 1), (4, "batched", 296, 1), (4, "batched", 444, 1), (8, "batched", 296, 1), (8, "groups:2", 148, 1),
           (8, "groups:2", 296, 1), (8, "groups:2", 444, 1), (8, "groups:3", 296, 1), (16, "groups:2", 296, 1), (8, "groups:2", 296, 0)]
if a.configs:
    CONFIGS = [tuple(int(x) if x.lstrip("-").isdigit() else x for x in c.split(",")) for c in a.configs.split(";")]


def run(cfg):
    chunks, inverse, ctas, prio = cfg[:4]
    kind = cfg[4] if len(cfg) > 4 else "peer"
    pipe.exchange_kind = kind
    if kind == "push":
        pipe.push_ctas = ctas
        ctas = 0
    pipe.chunks = chunks
    pipe.inverse_batched = inverse == "batched"
    pipe.inverse_groups = int(inverse.split(":")[1]) if inverse.startswith("groups:") else 0
    L.set_option("peer_pass_ctas", ctas)
    pipe._side = torch.cuda.Stream(priority=-1 if prio else 0)
    for _ in range(2):
        ti.do_advance(data, dt)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        ti.do_advance(data, dt)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = []
for cfg in CONFIGS:
    ms = run(cfg)
    if rank == 0:
        rec = {"world": world, "n": a.n, "chunks": cfg[0], "inverse": cfg[1], "peer_pass_ctas": cfg[2], "side_priority": cfg[3],
               "exchange": cfg[4] if len(cfg) > 4 else "peer", "ms_per_step": round(ms, 3), "upd_per_s": 4 * nk / (ms * 1e-3)}
        out.append(rec)
        print(json.dumps(rec), flush=True)
# per-kernel CUDA-event totals of one step for the first and the best configuration (launch labels of csrc/; "push" = the copy kernel)
best = min(range(len(CONFIGS)), key=lambda i: out[i]["ms_per_step"]) if rank == 0 else 0
bt = torch.tensor([best], device="cuda")
dist.broadcast(bt, 0)
for i in sorted({0, int(bt.item())}):
    run(CONFIGS[i])
    L.profile(True)
    ti.do_advance(data, dt)
    prof = L.profile_report()
    L.profile(False)
    if rank == 0:
        print(json.dumps({"world": world, "config": list(CONFIGS[i]), "kernel_ms_per_step": {k: round(v["ms"], 3) for k, v in prof.items()},
                          "launches": {k: v["n"] for k, v in prof.items()}}), flush=True)
ek = va.ekin(data, reduce_all=True)
if rank == 0:
    print(json.dumps({"world": world, "ekin_after_all": ek, "best": min(out, key=lambda r: r["ms_per_step"])}))
dist.destroy_process_group()
