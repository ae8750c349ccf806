"""Timeline of the device-chained Lanczos step from the in-kernel event trace (b2k_debug_trace): runs one
b2k_lanczos_expand_many batch of the headline job's shape (n = 1e7 in total, CGS2, basis 36 -> 60) on 1 GPU or under
torchrun on N GPUs, and prints, per rank, the mean time between consecutive events of a step.

    python tools/trace_step.py                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29777 tools/trace_step.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krylovkit_jl_b200 as kk  # noqa: E402
from krylovkit_jl_b200 import sharding  # noqa: E402
from krylovkit_jl_b200.factorizations import lanczos as lz  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local_rank = int(os.environ.get("LOCAL_RANK", 0))
dist = None
lib = kk._lib.load()
nx, ny, kd = 4000, 2500, 60
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    uid = sharding.broadcast_nccl_uid(dist, lib)
    sh = sharding.shard_grid_lines(nx, ny, rank, world)
    ctx = kk.B200Context(sh.n_local, kd + 12, device=local_rank, rank=rank, nranks=world, nccl_uid=uid,
                         n_global=nx * ny, row_offset=sh.row_offset)
else:
    ctx = kk.B200Context(nx * ny, kd + 12, device=local_rank)
op = kk.B200CSR.stencil(ctx, nx, ny)
x0 = ctx.splitmix(20240607)
it = lz.LanczosIterator(op, x0, kk.cgs2)
f = lz.initialize(it)
lz.expand_many_(it, f, 35, 0.0)                    # basis 1 -> 36 (warm-up, not traced)
ctx.check(lib.b2k_debug_trace(ctx.h, 1))
lz.expand_many_(it, f, 24, 0.0)                    # 36 -> 60: the traced batch
cap = 1 << 16
buf = (C.c_uint64 * (2 * cap))()
n = C.c_int64()
ctx.check(lib.b2k_debug_trace_read(ctx.h, buf, cap, C.byref(n)))
ev = np.frombuffer(buf, dtype=np.uint64)[: 2 * n.value].reshape(-1, 2).astype(np.int64)
ev = ev[np.argsort(ev[:, 0], kind="stable")]
names = {1: "spmv begin", 2: "spmv halo rows present", 3: "spmv CTA0 tiles done", 4: "spmv <v,Av> published (last CTA)",
         10: "sweep begin", 11: "sweep alpha present", 12: "project phase done (CTA0)", 15: "boundary left",
         13: "update phase done (CTA0)", 18: "finaliser entered (last CTA)", 19: "finaliser left"}
# the Lanczos CGS2 step (lanczos.jl:313-324) is two phases: [three-term prologue + projection] | [update + norm]
order = [1, 2, 3, 4, 10, 11, 12, 15, 13, 18, 19]
# split into steps at every "spmv begin"
starts = np.flatnonzero(ev[:, 1] == 1)
steps = []
for a, b in zip(starts, list(starts[1:]) + [len(ev)]):
    d = {int(c): int(t) for t, c in ev[a:b]}
    if all(c in d for c in order):
        steps.append((d, int(ev[b, 0]) if b < len(ev) else None))
rows = []
for i in range(len(order) - 1):
    dt = [s[order[i + 1]] - s[order[i]] for s, _ in steps]
    rows.append((names[order[i]] + " -> " + names[order[i + 1]], float(np.mean(dt)) / 1e3, float(np.max(dt)) / 1e3))
gap = [nxt - s[19] for s, nxt in steps if nxt is not None]
rows.append(("finaliser left -> next spmv begin (kernel boundary)", float(np.mean(gap)) / 1e3 if gap else 0.0,
             float(np.max(gap)) / 1e3 if gap else 0.0))
tot = [nxt - s[1] for s, nxt in steps if nxt is not None]
out = {"rank": rank, "world": world, "steps": len(steps), "us_per_step": float(np.mean(tot)) / 1e3 if tot else None,
       "intervals_us(mean,max)": {k: (round(m, 2), round(x, 2)) for k, m, x in rows}}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"trace_step_n{world}_r{rank}.json"), "w"), indent=1)
if rank in (0, world // 2):
    print(f"--- rank {rank} of {world}: {len(steps)} steps, {out['us_per_step']:.1f} us per step (basis 36..60)")
    for k, m, x in rows:
        print(f"  {k:62s} {m:8.2f} us  (max {x:8.2f})")
if dist is not None:
    dist.barrier()
ctx.close()
