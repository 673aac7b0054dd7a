"""N>1 path on CPU: world_size-2 (and 3) gloo process groups run the SAME host logic as the GPUs
(sparseframe_b200.parallel.PartitionedFactorization: exchange of the ranks' handles, connect, factorize) over the CPU plan
interpreter of the oracle.  The interpreters' images are shared-memory segments -- the stand-in for CUDA IPC mappings --
so remote update entries and pushed blocks really cross process boundaries, and the union of the owned panels must
reproduce the golden factor."""
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


def _worker(rank, world, port, gen, g, kind, devslot, group_panels, out_path):
    sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import sforacle
    import sparseframe_b200 as sf
    from sparseframe_b200.parallel import PartitionedFactorization
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SFCHECK_DIST_MIN_M"] = "8"       # tiny fronts too: exercise the push + tile-share path
    # group_panels: shared Cholesky fronts of >= 24 columns are factorized by their whole group in 16-wide block columns
    os.environ["SFCHECK_DIST_PANEL_MIN"] = "24" if group_panels else "1000000"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = sf.compress(*sf.matrices.generate(gen, g))
    backend = sforacle.CheckBackend(kind, A, devslot, rank, world, nb=8 if group_panels else 0, outer=16 if group_panels else 0,
                                    dist=dist)
    pf = PartitionedFactorization(backend, dist)
    for _ in range(2):                                  # a second factorization re-zeroes and re-uses the mappings
        pf.factorize()
        assert backend.bad == 0
    owned = torch.from_numpy(backend.pack_owned())      # zeros outside the owned panels -> SUM = union
    dist.reduce(owned, dst=0, op=dist.ReduceOp.SUM)
    ntypes = np.bincount([c["type"] for c in backend.colls], minlength=2)
    stats = torch.tensor([backend.double("local_flops"), float(backend.pushed_bytes), float(backend.int("xsize_pad")),
                          float(backend.image_elems[rank])], dtype=torch.float64)
    gathered = [torch.zeros(4, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    dist.gather(stats, gathered, dst=0)
    if rank == 0:
        np.savez(out_path, lsx=owned.numpy(), stats=torch.stack(gathered).numpy(), total=backend.double("flops"),
                 ncoll=ntypes, full_image=float(sum(int(c) for c in [0])))
    dist.barrier()
    dist.destroy_process_group()
    backend.close()


@pytest.mark.parametrize("world,name,group_panels", [
    (2, "lap3d_10_chol_slot", 0), (3, "lap3d_10_chol_slot", 0), (2, "cd3d_8_lu", 0), (3, "cd3d_8_lu", 0),
    (2, "lap2d_20_chol", 0), (3, "lap2d_20_chol", 0),
    (2, "lap3d_10_chol_slot", 1), (3, "lap3d_10_chol_slot", 1), (3, "cd3d_8_lu", 1)])
def test_partitioned_factorization_over_gloo(world, name, group_panels, tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden
    import sparseframe_b200 as sf
    gold = load_golden(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    out = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(world, _free_port(), gold["gen"], gold["grid"], gold["kind"], gold["devslot"], group_panels, out),
             nprocs=world, join=True)
    res = np.load(out)
    mask = sf.lower_trapezoid_mask(gold["kind"], gold["Super"], gold["Lsip"], gold["Lsxp"], gold["xsize"])
    a, b = gold["Lsx"][mask], res["lsx"][mask]
    assert sf.entrywise_rel_diff(a, b) <= 1e-10
    assert np.array_equal(a == 0, b == 0)
    # every flop is owned by exactly one rank; there is at least the start barrier and one group barrier
    ncoll = res["ncoll"]
    assert res["stats"][:, 0].sum() == pytest.approx(float(res["total"]), rel=1e-12)
    assert ncoll[0] >= 2
    # shared fronts with off-diagonal rows are pushed to their group and their Schur update is split over it
    assert (res["stats"][:, 1].sum() > 0) == (ncoll[1] > 0)
    if name == "lap3d_10_chol_slot":
        assert ncoll[1] > 0, "this case must exercise the distributed Schur updates"
    # every rank computed every rank's layout and they agree; no rank holds the whole factor any more
    assert np.array_equal(res["stats"][:, 2], res["stats"][:, 3])


@pytest.mark.parametrize("group_panels", [0, 1])
def test_partition_simulated_up_to_8_ranks(group_panels, monkeypatch):
    """All ranks simulated inside one process (oracle/plan_host_exec.cpp): ownership, collective lists and segments
    stay consistent for world sizes up to 8, and the result equals the compiled reference / oracle.  group_panels: the
    wide shared Cholesky fronts are factorized by their groups (16-wide block columns owned round-robin)."""
    monkeypatch.setenv("SFCHECK_DIST_PANEL_MIN", "32" if group_panels else "1000000")
    monkeypatch.setenv("SFCHECK_DIST_MIN_M", "24" if group_panels else "512")
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sforacle
    import sparseframe_b200 as sf
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsfcheck.so"))
    lib.sfcheck_factorize_partitioned.restype = C.c_long
    lib.sfcheck_factorize_partitioned.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_ulonglong] + [C.c_int] * 5 + [C.c_void_p] * 4
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
        plan_elems = None
        for world in (1, 2, 4, 5, 8):
            for reversed_ranks in (False, True):        # the result must not depend on which rank of a segment runs first
                if reversed_ranks:
                    monkeypatch.setenv("SFCHECK_RANKS_REVERSED", "1")
                else:
                    monkeypatch.delenv("SFCHECK_RANKS_REVERSED", raising=False)
                out = np.full(S.int("xsize"), np.nan)
                bal = np.zeros(world)
                ncoll = (C.c_long * 2)()
                image = (C.c_long * world)()
                rc = lib.sfcheck_factorize_partitioned(int(kind == "lu"), int(A.symmetric), A.n, A.Cp.ctypes.data,
                                                       A.Ci.ctypes.data, A.Cx.ctypes.data, sf.DEFAULT_DEVSLOT,
                                                       8 if group_panels else 0, 16 if group_panels else 0, -1, -1,
                                                       world, out.ctypes.data, bal.ctypes.data, ncoll, image)
                assert rc == 0
                assert sf.entrywise_rel_diff(ref[mask], out[mask]) <= 1e-10
                assert bal.sum() == pytest.approx(S.flops, rel=1e-12)
                if world == 1:
                    plan_elems = image[0]
                    assert ncoll[0] == 0 and ncoll[1] == 0
                    continue
                assert ncoll[0] >= 2                         # the start barrier and at least one group barrier
                assert bal.max() / bal.sum() < 0.75          # no rank is left with (almost) everything
                # images are compact: a rank stores its own panels and the shared fronts of its groups, not the whole factor
                assert max(image) < plan_elems and sum(image) >= plan_elems


@pytest.mark.parametrize("gen,g,kind,group_panels", [("lap3d", 14, "cholesky", 0), ("lap3d", 14, "cholesky", 1), ("cd3d", 12, "lu", 0),
                                                     ("lap2d", 60, "cholesky", 1)])
def test_exchange_protocol_under_real_concurrency(gen, g, kind, group_panels, monkeypatch):
    """The flags themselves: one host thread per rank, NO lock step, three factorizations back to back, the plan's own
    sequence numbers and epochs (oracle/plan_host_exec.cpp: sfcheck_factorize_threaded).  A remote add that is not fenced by
    a barrier, a wait nobody signals, or a start barrier that lets a fast rank add into an image its owner has not zeroed
    yet shows up as a wrong factor or a timeout; per-step jitter varies the interleaving from run to run."""
    monkeypatch.setenv("SFCHECK_DIST_PANEL_MIN", "32" if group_panels else "1000000")
    monkeypatch.setenv("SFCHECK_DIST_MIN_M", "24" if group_panels else "48")
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sforacle
    import sparseframe_b200 as sf
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsfcheck.so"))
    lib.sfcheck_factorize_threaded.restype = C.c_long
    lib.sfcheck_factorize_threaded.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong] + \
        [C.c_int] * 7 + [C.c_double, C.c_void_p]
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
    for world in (2, 3, 4, 8):
        for jitter in (0, 40):
            out = np.full(S.int("xsize"), np.nan)
            rc = lib.sfcheck_factorize_threaded(int(kind == "lu"), int(A.symmetric), A.n, A.Cp.ctypes.data, A.Ci.ctypes.data,
                                                A.Cx.ctypes.data, sf.DEFAULT_DEVSLOT, 8 if group_panels else 0,
                                                16 if group_panels else 0, -1, -1, world, 3, jitter, 60.0, out.ctypes.data)
            assert rc == 0, "world %d jitter %d: rc %d (-300 = a flag wait timed out)" % (world, jitter, rc)
            assert not np.isnan(out[mask]).any()
            assert sf.entrywise_rel_diff(ref[mask], out[mask]) <= 1e-10, (world, jitter)
            assert np.array_equal(ref[mask] == 0, out[mask] == 0)


def test_concurrency_checker_notices_missing_waits(monkeypatch):
    """Self-test of the threaded checker: with every flag wait dropped the ranks run past each other and the factor is wrong."""
    monkeypatch.setenv("SFCHECK_DIST_PANEL_MIN", "32")
    monkeypatch.setenv("SFCHECK_DIST_MIN_M", "24")
    monkeypatch.setenv("SFCHECK_THREADED_NO_WAIT", "1")
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sforacle
    import sparseframe_b200 as sf
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsfcheck.so"))
    lib.sfcheck_factorize_threaded.restype = C.c_long
    lib.sfcheck_factorize_threaded.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong] + \
        [C.c_int] * 7 + [C.c_double, C.c_void_p]
    A = sf.compress(*sf.matrices.generate("lap3d", 14))
    S = sf.analyze(A, "cholesky")
    st = {k: S.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
    ref, _ = sforacle.factorize("cholesky", st, S.array("Lx"), None)
    mask = sf.lower_trapezoid_mask("cholesky", st["Super"], st["Lsip"], st["Lsxp"], S.int("xsize"))
    wrong = 0
    for jitter in (0, 20, 60, 100):
        out = np.full(S.int("xsize"), np.nan)
        rc = lib.sfcheck_factorize_threaded(0, 1, A.n, A.Cp.ctypes.data, A.Ci.ctypes.data, A.Cx.ctypes.data, sf.DEFAULT_DEVSLOT,
                                            8, 16, -1, -1, 4, 3, jitter, 30.0, out.ctypes.data)
        good = rc == 0 and not np.isnan(out[mask]).any() and sf.entrywise_rel_diff(ref[mask], out[mask]) <= 1e-10
        wrong += not good
    assert wrong > 0
