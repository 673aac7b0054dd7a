"""Multi-GPU check (run under torchrun, one rank per GPU): etree-partitioned factorization with NCCL reduces, union
of the owned panels compared with the single-GPU factor, then timing.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/mgpu_check.py lap3d 40
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
import sparseframe_b200 as sf  # noqa: E402
from sparseframe_b200.parallel import DeviceBackend, PartitionedFactorization, owned_mask  # noqa: E402

gen, g = sys.argv[1], int(sys.argv[2])
kind = sys.argv[3] if len(sys.argv) > 3 else "cholesky"
check = len(sys.argv) <= 4 or sys.argv[4] != "nocheck"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t0 = time.time()
A = sf.compress(*sf.matrices.generate(gen, g))
S = sf.analyze(A, kind)
t_an = time.time() - t0
split = kind == "lu" and S.int("symmetric") == 0
plan = sf.Plan(S, device=local, rank=rank, world=world)
stream = torch.cuda.Stream()          # capturable (the legacy default stream is not): segments replay as CUDA graphs
torch.cuda.set_stream(stream)
plan.set_stream(stream.cuda_stream)
d_lx = torch.from_numpy(S.array("Lx")).cuda()
d_ux = torch.from_numpy(S.array("Ux")).cuda() if split else None
pf = PartitionedFactorization(DeviceBackend(plan, d_lx, d_ux), dist)
share = plan.double("local_flops") / S.flops
for it in range(4):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    pf.factorize()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("iter %d: %.2f ms  %.2f TFLOP/s (world %d)" % (it, ms.item(), S.flops / ms.item() / 1e9, world), flush=True)
print("rank %d: analyze %.1fs, owns %.3f of the flops, collectives %d, reduced %.3f GB, segments %d, launches %d" % (
    rank, t_an, share, plan.int("ncolls"), pf.reduced_bytes / 1e9, plan.int("nsegments"), plan.int("nlaunches")), flush=True)
if check:
    mine = plan.download()
    mask = owned_mask(plan.owner(), S.array("Lsxp"), rank)
    mine[~mask] = 0.0
    t = torch.from_numpy(mine).cuda()
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        merged = t.cpu().numpy()
        single = sf.Plan(S, device=local)
        ref = single.factorize(S.array("Lx"), S.array("Ux") if split else None)
        m = sf.lower_trapezoid_mask(kind, S.array("Super"), S.array("Lsip"), S.array("Lsxp"), S.int("xsize"))
        print("multi-GPU vs single-GPU factor: max rel diff %.2e, zero-pattern equal %s, single-GPU %.2f ms" % (
            sf.entrywise_rel_diff(ref[m], merged[m]), bool(np.array_equal(ref[m] == 0, merged[m] == 0)), single.last_ms(0)),
            flush=True)
dist.barrier()
dist.destroy_process_group()
