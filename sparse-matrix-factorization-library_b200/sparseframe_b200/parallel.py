"""Etree-subtree-partitioned factorization over several GPUs of one box: one process per GPU,
torch.distributed for the plumbing.

Each rank factorizes the subtrees it owns.  Updates aimed at a "top" supernode owned by another rank are accumulated in
the rank's own zero-initialised image of that panel; before the owner factorizes the level of that supernode, the
images are summed into the owner with a reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).  That sum is the one
real exchange step of the path -- the reference's equivalent is its host-RAM round trip of factored panels
(CH.c:2596-2685); nothing else is communicated.

The phase loop is backend-agnostic so that the same orchestration code runs on the GPU engine (`DeviceBackend`) and, in
the CPU test tier, on the plan interpreter of the oracle (tests/test_multirank_gloo.py).
"""
from __future__ import annotations

import numpy as np


class _CudaView:
    """Zero-copy torch view of a raw device allocation through __cuda_array_interface__."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


class DeviceBackend:
    """One rank's sfb200_plan (created with rank/world) behind the backend protocol."""

    def __init__(self, plan, d_lx, d_ux=None):
        import torch
        self.plan = plan
        self.rank, self.world = plan.rank, plan.world
        self.nphases = plan.int("nphases")
        self.d_lx, self.d_ux = d_lx, d_ux
        self.image = torch.as_tensor(_CudaView(plan.factor_ptr(), plan.int("xsize_padded")), device=d_lx.device)
        self.phases = [plan.phase_info(k) for k in range(self.nphases)]
        self.reduces = [plan.reduce_info(i) for i in range(plan.int("nreduces"))]
        self.bcasts = [plan.bcast_info(i) for i in range(plan.int("nbcasts"))]

    def begin(self):
        self.plan.phase_begin(self.d_lx.data_ptr(), self.d_ux.data_ptr() if self.d_ux is not None else 0)

    def run_phase(self, k, part=0):
        self.plan.phase_run(k, part)

    def finish(self):
        self.plan.sync()


class PartitionedFactorization:
    """Phase loop, one phase per etree level: reduce the partial images of the level's shared panels into their owners,
    run the level's launches (A), broadcast the shared panels that were just factorized to their groups, run every rank's
    share of their Schur-update tiles (B)."""

    def __init__(self, backend, dist=None):
        self.b = backend
        self.dist = dist
        self.groups = {}
        if dist is not None and backend.world > 1:
            # groups must be created by every rank, in the same order: the reduce list is identical on all ranks
            for r in list(backend.reduces) + list(backend.bcasts):
                key = (r["lo"], r["hi"])
                if key not in self.groups:
                    full = key == (0, backend.world)
                    self.groups[key] = None if full else dist.new_group(ranks=list(range(key[0], key[1])))
        self.reduced_bytes = 0
        self.bcast_bytes = 0

    def factorize(self):
        b = self.b
        b.begin()
        self.reduced_bytes = 0
        self.bcast_bytes = 0
        for k in range(b.nphases):
            ph = b.phases[k]
            for i in range(ph["reduce_first"], ph["reduce_first"] + ph["reduce_count"]):
                r = b.reduces[i]
                if b.world > 1 and r["lo"] <= b.rank < r["hi"]:
                    view = b.image[r["offset"]:r["offset"] + r["count"]]
                    self.dist.reduce(view, dst=r["root"], op=self.dist.ReduceOp.SUM, group=self.groups[(r["lo"], r["hi"])])
                    self.reduced_bytes += 8 * r["count"]
            b.run_phase(k, 0)
            # factored shared panels go from their owner to the group, then everybody computes its share of their
            # Schur-update tiles (part B of the phase)
            for i in range(ph["bcast_first"], ph["bcast_first"] + ph["bcast_count"]):
                r = b.bcasts[i]
                if b.world > 1 and r["lo"] <= b.rank < r["hi"]:
                    view = b.image[r["offset"]:r["offset"] + r["count"]]
                    self.dist.broadcast(view, src=r["root"], group=self.groups[(r["lo"], r["hi"])])
                    self.bcast_bytes += 8 * r["count"]
            if ph["launch_b_count"] > 0:
                b.run_phase(k, 1)
        b.finish()


def owned_mask(owner, Lsxp, rank):
    """Boolean mask over the reference's Lsx layout of the entries whose panel `rank` owns."""
    mask = np.zeros(int(Lsxp[-1]), dtype=bool)
    for s in np.flatnonzero(np.asarray(owner) == rank):
        mask[Lsxp[s]:Lsxp[s + 1]] = True
    return mask
