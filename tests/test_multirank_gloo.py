"""N>1 path on CPU: world_size-2 (and 3) gloo process groups run the SAME orchestration code as the GPUs
(sparseframe_b200.parallel.PartitionedFactorization) over the CPU plan interpreter of the oracle, and the union of the
owned panels must reproduce the golden factor."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, gen, g, kind, devslot, out_path):
    sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import sforacle
    import sparseframe_b200 as sf
    from sparseframe_b200.parallel import PartitionedFactorization
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SFCHECK_DIST_MIN_M"] = "8"       # tiny fronts too: exercise the broadcast + tile-share path
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = sf.compress(*sf.matrices.generate(gen, g))
    backend = sforacle.CheckBackend(kind, A, devslot, rank, world)
    pf = PartitionedFactorization(backend, dist)
    pf.factorize()
    assert backend.bad == 0
    owned = torch.from_numpy(backend.pack_owned())      # zeros outside the owned panels -> SUM = union
    dist.reduce(owned, dst=0, op=dist.ReduceOp.SUM)
    stats = torch.tensor([backend.double("local_flops"), float(pf.reduced_bytes), float(pf.bcast_bytes)], dtype=torch.float64)
    gathered = [torch.zeros(3, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    dist.gather(stats, gathered, dst=0)
    if rank == 0:
        np.savez(out_path, lsx=owned.numpy(), stats=torch.stack(gathered).numpy(), total=backend.double("flops"),
                 nreduces=backend.int("nreduces"), nbcasts=backend.int("nbcasts"))
    dist.barrier()
    dist.destroy_process_group()
    backend.close()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["lap3d_10_chol_slot", "cd3d_8_lu", "lap2d_20_chol"])
def test_partitioned_factorization_over_gloo(world, name, tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden
    import sparseframe_b200 as sf
    gold = load_golden(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    out = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(world, _free_port(), gold["gen"], gold["grid"], gold["kind"], gold["devslot"], out),
             nprocs=world, join=True)
    res = np.load(out)
    mask = sf.lower_trapezoid_mask(gold["kind"], gold["Super"], gold["Lsip"], gold["Lsxp"], gold["xsize"])
    a, b = gold["Lsx"][mask], res["lsx"][mask]
    assert sf.entrywise_rel_diff(a, b) <= 1e-10
    assert np.array_equal(a == 0, b == 0)
    # every flop is owned by exactly one rank, and something really had to be exchanged
    assert res["stats"][:, 0].sum() == pytest.approx(float(res["total"]), rel=1e-12)
    assert int(res["nreduces"]) > 0 and res["stats"][:, 1].sum() > 0
    # shared fronts with off-diagonal rows are broadcast and their Schur update is split over the group
    assert (res["stats"][:, 2].sum() > 0) == (int(res["nbcasts"]) > 0)
    if name == "lap3d_10_chol_slot":
        assert int(res["nbcasts"]) > 0, "this case must exercise the distributed Schur updates"


def test_partition_simulated_up_to_8_ranks():
    """All ranks simulated inside one process (oracle/plan_host_exec.cpp): ownership, reduce lists and phases stay
    consistent for world sizes up to 8, and the result equals the compiled reference / oracle."""
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sforacle
    import sparseframe_b200 as sf
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsfcheck.so"))
    lib.sfcheck_factorize_partitioned.restype = C.c_long
    lib.sfcheck_factorize_partitioned.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_ulonglong] + [C.c_int] * 5 + [C.c_void_p] * 3
    for gen, g, kind in (("lap3d", 14, "cholesky"), ("cd3d", 12, "lu")):
        A = sf.compress(*sf.matrices.generate(gen, g))
        S = sf.analyze(A, kind)
        st = {k: S.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
        st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
        ux = None
        if kind == "lu":
            st["Up"], st["Ui"] = S.array("Up"), S.array("Ui")
            ux = S.array("Ux")
        ref, bad = sforacle.factorize(kind, st, S.array("Lx"), ux)
        assert bad == 0
        mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], S.int("xsize"))
        for world in (2, 4, 5, 8):
            out = np.full(S.int("xsize"), np.nan)
            bal = np.zeros(world)
            nred = C.c_long()
            rc = lib.sfcheck_factorize_partitioned(int(kind == "lu"), int(A.symmetric), A.n, A.Cp.ctypes.data,
                                                   A.Ci.ctypes.data, A.Cx.ctypes.data, sf.DEFAULT_DEVSLOT, 0, 0, -1, -1,
                                                   world, out.ctypes.data, bal.ctypes.data, C.byref(nred))
            assert rc == 0
            assert sf.entrywise_rel_diff(ref[mask], out[mask]) <= 1e-10
            assert bal.sum() == pytest.approx(S.flops, rel=1e-12)
            assert bal.max() / bal.sum() < 0.75          # no rank is left with (almost) everything
