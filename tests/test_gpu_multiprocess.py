"""One PROCESS per rank, as bench.py runs under torchrun: every rank builds its part of the partitioned plan, the ranks
exchange CUDA IPC handles through torch.distributed (sparseframe_b200.parallel), map each other's images and flags, and
factorize -- remote adds, pushes and flag waits then cross process boundaries.  With at least `world` GPUs every rank gets
its own device and the handle exchange runs over NCCL; on a smaller box the ranks share devices (time-sliced contexts) and
the handles travel over gloo.  The union of the owned panels must equal the oracle's factor."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, gen, g, kind, env, ngpu, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
    os.environ.update(env)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import sparseframe_b200 as sf
    from sparseframe_b200.parallel import DeviceBackend, PartitionedFactorization
    device = rank % ngpu
    torch.cuda.set_device(device)
    own_gpu = ngpu >= world
    if own_gpu:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    A = sf.compress(*sf.matrices.generate(gen, g))
    S = sf.analyze(A, kind)
    split = kind == "lu" and S.int("symmetric") == 0
    plan = sf.Plan(S, device=device, rank=rank, world=world)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        d_lx = torch.from_numpy(S.array("Lx")).cuda()
        d_ux = torch.from_numpy(S.array("Ux")).cuda() if split else None
        pf = PartitionedFactorization(DeviceBackend(plan, d_lx, d_ux), dist)     # binds the plan to `stream`, connects the ranks
        assert plan.int("connected") == 1
        for it in range(3):             # plain, graph capture, replay
            if it == 2:
                d_lx.mul_(4.0)          # new values through the captured graph: chol(4 A) = 2 chol(A)
            pf.factorize()
    out = np.full(S.int("xsize"), np.nan)
    plan.download(out)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), out)
    np.save(os.path.join(out_dir, "owner%d.npy" % rank), plan.owner())
    dist.barrier()
    dist.destroy_process_group()
    plan.close()


@pytest.mark.parametrize("gen,g,kind,world,group_panels", [("lap3d", 24, "cholesky", 2, 1), ("lap3d", 20, "cholesky", 3, 0)])
def test_ranks_in_separate_processes_connect_over_cuda_ipc(gen, g, kind, world, group_panels, tmp_path):
    import torch
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sforacle
    import sparseframe_b200 as sf
    from sparseframe_b200.parallel import owned_mask
    ngpu = torch.cuda.device_count()
    env = dict(SFB200_DIST_MIN_M="48", SFB200_DIST_PANEL_MIN="256" if group_panels else "1000000", SFB200_PEER_TIMEOUT_S="60")
    if group_panels:
        env.update(SFB200_OUTER="128", SFB200_OUTER2="256")
    mp.spawn(_worker, args=(world, _free_port(), gen, g, kind, env, ngpu, str(tmp_path)), nprocs=world, join=True)
    A = sf.compress(*sf.matrices.generate(gen, g))
    S = sf.analyze(A, kind)
    st = {k: S.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
    ref, bad = sforacle.factorize(kind, st, 4.0 * S.array("Lx"), None)
    assert bad == 0
    owner = np.load(tmp_path / "owner0.npy")
    merged = np.full(S.int("xsize"), np.nan)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("owner%d.npy" % r)), owner)
        m = owned_mask(owner, st["Lsxp"], r)
        part = np.load(tmp_path / ("rank%d.npy" % r))
        assert np.isnan(part[~m]).all(), "a rank wrote panels it does not own"
        merged[m] = part[m]
    mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], len(ref))
    assert not np.isnan(merged[mask]).any()
    assert sf.entrywise_rel_diff(ref[mask], merged[mask]) <= 1e-10
    assert np.array_equal(ref[mask] == 0, merged[mask] == 0)
