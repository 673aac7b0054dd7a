"""Etree-subtree-partitioned factorization over several GPUs of one box, one process per GPU.

Each rank factorizes the subtrees it owns.  The supernodes above them ("top") are shared by a group of ranks; the
exchange medium is peer memory over NVLink (include/sfb200.h, csrc/plan_types.h `Coll`):
  * an update aimed at a panel another rank stores is added straight into that rank's image by the scatter epilogue of
    the update kernels (`red.global.add.f64` through the peer mapping);
  * wide shared Cholesky fronts are factorized by their whole group (block columns owned round-robin), the Schur update
    of a shared front is split over its group; factored blocks are pushed to the group with peer-to-peer copies;
  * synchronisation is flags in peer memory: no host round trip, a rank's whole factorization replays as one CUDA graph.
The reference's equivalent is its host-RAM round trip of factored panels (CH.c:2596-2685).

What is left for the host layer -- and all this module does -- is to CONNECT the ranks once: every rank exports a
128-byte handle of its image and flag array (CUDA IPC), the handles travel through torch.distributed (NCCL or gloo: plumbing,
no payload ever goes through it), every rank maps its peers.  The same code connects the CPU plan interpreters of the test
tier, whose "images" are shared-memory segments (tests/test_multirank_gloo.py).
"""
from __future__ import annotations

import numpy as np

HANDLE_BYTES = 128


class _CudaView:
    """Zero-copy torch view of a raw device allocation through __cuda_array_interface__."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


def all_gather_handles(dist, handle: bytes, world: int) -> bytes:
    """The ranks' exchange handles, concatenated in rank order."""
    import torch
    if len(handle) != HANDLE_BYTES:
        raise ValueError("an exchange handle has %d bytes" % HANDLE_BYTES)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(dev)
    parts = [torch.empty(HANDLE_BYTES, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in parts)


class DeviceBackend:
    """One rank's sfb200_plan (created with rank / world) and its matrix values in HBM."""

    def __init__(self, plan, d_lx, d_ux=None):
        import torch
        self.plan = plan
        self.rank, self.world = plan.rank, plan.world
        self.d_lx, self.d_ux = d_lx, d_ux
        # The plan's launches must be ordered after whatever produced d_lx / d_ux and before whatever reads the factor:
        # both live on torch's current stream, so the plan runs there (a plan left on its private stream would race with them).
        self.stream = torch.cuda.current_stream(d_lx.device)
        plan.set_stream(self.stream.cuda_stream)

    def exchange_handle(self) -> bytes:
        return self.plan.exchange_handle()

    def connect(self, handles: bytes):
        self.plan.connect(handles)

    def factorize(self):
        """This rank's part: asynchronous on the plan's stream; every rank of the partition must call it."""
        self.plan.factorize_device(self.d_lx.data_ptr(), self.d_ux.data_ptr() if self.d_ux is not None else 0)

    def finish(self):
        self.plan.sync()


class PartitionedFactorization:
    """Connects the ranks (once), then every factorize() is one device-side replay per rank."""

    def __init__(self, backend, dist):
        self.b = backend
        self.dist = dist
        if backend.world > 1:
            backend.connect(all_gather_handles(dist, backend.exchange_handle(), backend.world))

    def factorize(self, wait=True):
        self.b.factorize()
        if wait:
            self.b.finish()


def owned_runs(owner, my_offsets, global_offsets, rank):
    """Runs of consecutive supernodes `rank` owns: (offset in its image, offset in the single-GPU image, elements).
    Consecutive stored panels are back to back in both layouts."""
    owner = np.asarray(owner)
    mine = np.flatnonzero(owner == rank)
    runs = []
    if mine.size == 0:
        return runs
    brk = np.flatnonzero(np.diff(mine) != 1)
    starts = np.concatenate([[mine[0]], mine[brk + 1]])
    ends = np.concatenate([mine[brk], [mine[-1]]])
    for a, b in zip(starts, ends):
        count = int(global_offsets[b + 1] - global_offsets[a])
        runs.append((int(my_offsets[a]), int(global_offsets[a]), count))
    return runs


def owned_mask(owner, Lsxp, rank):
    """Boolean mask over the reference's Lsx layout of the entries whose panel `rank` owns."""
    mask = np.zeros(int(Lsxp[-1]), dtype=bool)
    for s in np.flatnonzero(np.asarray(owner) == rank):
        mask[Lsxp[s]:Lsxp[s + 1]] = True
    return mask
