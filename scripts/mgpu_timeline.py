"""Where does a partitioned step spend its time?  Same segment loop as sparseframe_b200.parallel.PartitionedFactorization,
with CUDA events around every collective and every segment's launches; per rank: time in reduces / all-reduces /
broadcasts / kernels, per level and in total, and the time the rank sat inside collectives (which includes waiting for the
slowest member of the group).  Measurement tool for DESIGN.md section 11, items 1-2 -- not a bench number (events between
the pieces keep them from overlapping exactly as in the untimed run).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/mgpu_timeline.py cfg5
"""
import collections
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
import sparseframe_b200 as sf  # noqa: E402
from sparseframe_b200.parallel import ALLREDUCE, BCAST, REDUCE, DeviceBackend, PartitionedFactorization  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
gen, g, kind = sf.matrices.CONFIGS[cfg]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
A = sf.compress(*sf.matrices.generate(gen, g))
S = sf.analyze(A, kind)
split = kind == "lu" and S.int("symmetric") == 0
plan = sf.Plan(S, device=local, rank=rank, world=world)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
plan.set_stream(stream.cuda_stream)
d_lx = torch.from_numpy(S.array("Lx")).cuda()
d_ux = torch.from_numpy(S.array("Ux")).cuda() if split else None
pf = PartitionedFactorization(DeviceBackend(plan, d_lx, d_ux), dist)
b = pf.b
for _ in range(3):                       # warm-up: plain issue, graph capture, replay
    pf.factorize()
torch.cuda.synchronize()
dist.barrier()

NAMES = {REDUCE: "reduce", ALLREDUCE: "all_reduce", BCAST: "broadcast"}
marks = []                               # (level, what, bytes, event before, event after)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream)
    return e


t_begin = ev()
b.begin()
for k, sg in enumerate(b.segments):
    for i in range(sg["coll_first"], sg["coll_first"] + sg["coll_count"]):
        c = b.colls[i]
        if not (c["lo"] <= rank < c["hi"]):
            continue
        view = b.image[c["offset"]:c["offset"] + c["count"]]
        group = pf.groups[(c["lo"], c["hi"])]
        e0 = ev()
        if c["type"] == REDUCE:
            dist.reduce(view, dst=c["root"], op=dist.ReduceOp.SUM, group=group)
        elif c["type"] == ALLREDUCE:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.broadcast(view, src=c["root"], group=group)
        marks.append((sg["level"], NAMES[c["type"]], 8 * c["count"], e0, ev()))
    e0 = ev()
    b.run_segment(k)
    marks.append((sg["level"], "kernels", 0, e0, ev()))
t_end = ev()
b.finish()
torch.cuda.synchronize()

total = t_begin.elapsed_time(t_end)
by_what = collections.defaultdict(lambda: [0.0, 0, 0])
by_level = collections.defaultdict(lambda: collections.defaultdict(float))
for lev, what, nbytes, e0, e1 in marks:
    ms = e0.elapsed_time(e1)
    by_what[what][0] += ms; by_what[what][1] += 1; by_what[what][2] += nbytes
    by_level[lev][what] += ms
lines = ["rank %d: step %.1f ms (instrumented)" % (rank, total)]
for what, (ms, n, nbytes) in sorted(by_what.items()):
    rate = " %6.1f GB/s payload" % (nbytes / ms / 1e6) if nbytes else ""
    lines.append("    %-10s %8.1f ms  %5.1f %%  %4d calls  %7.2f GB%s" % (what, ms, 100 * ms / total, n, nbytes / 1e9, rate))
if rank == 0 or os.environ.get("SFB200_TIMELINE_ALL"):
    lines.append("    per level (ms): " + "  ".join(
        "L%d[k %.0f r %.0f a %.0f b %.0f]" % (lev, d["kernels"], d["reduce"], d["all_reduce"], d["broadcast"])
        for lev, d in sorted(by_level.items())))
for r in range(world):                   # one rank at a time, in order
    if r == rank:
        print("\n".join(lines), flush=True)
    dist.barrier()
dist.destroy_process_group()
