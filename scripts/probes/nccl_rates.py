"""NCCL rates at the message sizes of the partitioned factorization (FP64 sum): reduce, all-reduce, broadcast over the
whole world and over the lower half (the plan's sub-groups).  Decision input for DESIGN.md section 11, items 1-2.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/probes/nccl_rates.py"""
import os

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
half = dist.new_group(ranks=list(range(max(world // 2, 1)))) if world > 2 else None
buf = torch.zeros(1 << 29, dtype=torch.float64, device="cuda")          # 4 GiB


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for group, name, members in ((None, "world", world), (half, "lower half", world // 2)):
    if name == "lower half" and (half is None or rank >= world // 2):
        continue
    for mb in (8, 52, 512, 4096):            # 52 MB = one 256-column block of a 25 600-row panel
        n = mb * (1 << 20) // 8
        v = buf[:n]
        for op, fn in (("reduce", lambda: dist.reduce(v, dst=0, op=dist.ReduceOp.SUM, group=group)),
                       ("all_reduce", lambda: dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)),
                       ("broadcast", lambda: dist.broadcast(v, src=0, group=group))):
            if name == "world":
                ms = timed(fn)
            else:                             # sub-group: time locally (dist.barrier / MAX would involve the other half)
                fn(); torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    fn()
                b.record(); torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 5
            if rank == 0:
                print("%-10s %-10s over %d ranks, %5d MB: %8.3f ms  %7.1f GB/s (payload / time)" % (
                    name, op, members, mb, ms, mb * (1 << 20) / ms / 1e6), flush=True)
dist.barrier()
dist.destroy_process_group()
