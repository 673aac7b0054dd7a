"""Etree-subtree-partitioned factorization over several GPUs of one box: one process per GPU,
torch.distributed for the plumbing.

Each rank factorizes the subtrees it owns.  The supernodes above them ("top") are shared by a group of ranks: every rank
accumulates its updates to them in its own zero-initialised image of the factor, and the plan lists the exchange steps
(csrc/plan_types.h `Coll`) that turn the images into the factor:
  reduce      the images of a shared panel are summed into its owner right before its level is factorized;
  all-reduce  ... into the whole group, for wide Cholesky panels that the group factorizes together (block columns
              owned round-robin, one broadcast per factorized block column);
  broadcast   a factored panel / block column goes to the group, which then splits the tiles of its Schur update.
NCCL over NVLink on GPUs, gloo in the CPU tests.  The reference's equivalent of these steps is its host-RAM round trip
of factored panels (CH.c:2596-2685); nothing else is communicated.

The segment loop is backend-agnostic so that the same orchestration code runs on the GPU engine (`DeviceBackend`) and,
in the CPU test tier, on the plan interpreter of the oracle (tests/test_multirank_gloo.py).
"""
from __future__ import annotations

import numpy as np


class _CudaView:
    """Zero-copy torch view of a raw device allocation through __cuda_array_interface__."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


REDUCE, ALLREDUCE, BCAST = 0, 1, 2


class DeviceBackend:
    """One rank's sfb200_plan (created with rank/world) behind the backend protocol."""

    def __init__(self, plan, d_lx, d_ux=None):
        import torch
        self.plan = plan
        self.rank, self.world = plan.rank, plan.world
        self.d_lx, self.d_ux = d_lx, d_ux
        self.image = torch.as_tensor(_CudaView(plan.factor_ptr(), plan.int("xsize_padded")), device=d_lx.device)
        self.segments = [plan.segment_info(k) for k in range(plan.int("nsegments"))]
        self.colls = [plan.coll_info(i) for i in range(plan.int("ncolls"))]

    def begin(self):
        self.plan.partition_begin(self.d_lx.data_ptr(), self.d_ux.data_ptr() if self.d_ux is not None else 0)

    def run_segment(self, k):
        self.plan.segment_run(k)

    def finish(self):
        self.plan.sync()


class PartitionedFactorization:
    """Segment loop: the segment's collectives on regions of the factor image, then its launches."""

    def __init__(self, backend, dist=None):
        self.b = backend
        self.dist = dist
        self.groups = {}
        if dist is not None and backend.world > 1:
            # groups must be created by every rank, in the same order: the collective list is identical on all ranks
            for c in backend.colls:
                key = (c["lo"], c["hi"])
                if key not in self.groups:
                    full = key == (0, backend.world)
                    self.groups[key] = None if full else dist.new_group(ranks=list(range(key[0], key[1])))
        self.bytes = {REDUCE: 0, ALLREDUCE: 0, BCAST: 0}

    @property
    def reduced_bytes(self):
        return self.bytes[REDUCE] + self.bytes[ALLREDUCE]

    @property
    def bcast_bytes(self):
        return self.bytes[BCAST]

    def factorize(self):
        b, dist = self.b, self.dist
        b.begin()
        self.bytes = {REDUCE: 0, ALLREDUCE: 0, BCAST: 0}
        for k, sg in enumerate(b.segments):
            for i in range(sg["coll_first"], sg["coll_first"] + sg["coll_count"]):
                c = b.colls[i]
                if b.world == 1 or not (c["lo"] <= b.rank < c["hi"]):
                    continue
                view = b.image[c["offset"]:c["offset"] + c["count"]]
                group = self.groups[(c["lo"], c["hi"])]
                if c["type"] == REDUCE:
                    dist.reduce(view, dst=c["root"], op=dist.ReduceOp.SUM, group=group)
                elif c["type"] == ALLREDUCE:
                    dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
                else:
                    dist.broadcast(view, src=c["root"], group=group)
                self.bytes[c["type"]] += 8 * c["count"]
            b.run_segment(k)
        b.finish()


def owned_mask(owner, Lsxp, rank):
    """Boolean mask over the reference's Lsx layout of the entries whose panel `rank` owns."""
    mask = np.zeros(int(Lsxp[-1]), dtype=bool)
    for s in np.flatnonzero(np.asarray(owner) == rank):
        mask[Lsxp[s]:Lsxp[s + 1]] = True
    return mask
