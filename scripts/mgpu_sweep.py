"""Partition-parameter sweep on N GPUs (torchrun): one analysis, then for every setting a fresh partitioned plan, 2 warm-up and 3
timed factorizations (device-resident, CUDA events, max over ranks).  Settings are environment variables the plan builder reads
at creation: SFB200_BALANCE_TOL, SFB200_WEIGHTED_BLOCKS, SFB200_DIST_MIN_M, SFB200_DIST_PANEL_MIN, SFB200_PUSH.
    torchrun --nproc-per-node 4 scripts/mgpu_sweep.py cfg5 "SFB200_BALANCE_TOL=1.05" "SFB200_BALANCE_TOL=1.02" ..."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
import sparseframe_b200 as sf  # noqa: E402
from sparseframe_b200.parallel import DeviceBackend, PartitionedFactorization  # noqa: E402

cfg = sys.argv[1]
settings = sys.argv[2:] or [""]
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gen, g, kind = sf.matrices.CONFIGS[cfg]
n, r, c, x, sym = sf.matrices.generate(gen, g)
A = sf.compress(n, r, c, x, sym)
order = torch.empty(n, dtype=torch.int64, device="cuda")
S = None
if rank == 0:
    S = sf.analyze(A, kind)
    order.copy_(torch.from_numpy(S.array("Order")))
dist.broadcast(order, src=0)
if rank != 0:
    S = sf.analyze(A, kind, order=order.cpu().numpy())
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
d_lx = torch.from_numpy(S.array("Lx")).cuda()
for setting in settings:
    keys = []
    for kv in setting.split():
        k, v = kv.split("=")
        os.environ[k] = v
        keys.append(k)
    t0 = time.time()
    plan = sf.Plan(S, device=local, rank=rank, world=world)
    pf = PartitionedFactorization(DeviceBackend(plan, d_lx, None), dist)
    for _ in range(2):
        pf.factorize(wait=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        pf.factorize(wait=False)
    e1.record(stream)
    plan.sync()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    info = torch.tensor([plan.int("xsize_padded") * 8e-9, plan.int("ncolls"), plan.double("remote_update_bytes") * 1e-9, plan.double("push_bytes") * 1e-9], dtype=torch.float64, device="cuda")
    allinfo = [torch.zeros_like(info) for _ in range(world)]
    dist.all_gather(allinfo, info)
    if rank == 0:
        a = torch.stack(allinfo).cpu().numpy()
        print("SWEEP %s N=%d [%s]: %.1f ms/step = %.1f TFLOP/s | exchange steps %d, image GB max %.1f, remote adds %.1f GB, pushes %.1f GB (setup %.1f s)"
              % (cfg, world, setting, float(ms.item()), S.flops / float(ms.item()) / 1e9, int(a[0, 1]), a[:, 0].max(), a[:, 2].sum(), a[:, 3].sum(), time.time() - t0), flush=True)
    plan.close()
    del pf
    for k in keys:
        os.environ.pop(k, None)
    dist.barrier()
dist.destroy_process_group()
