"""Parity of the CUDA path (through the C ABI, lib/libsfb200.so) against the oracle: golden vectors dumped from the
compiled reference, the plain-C restatement, the compiled reference itself when its .so travelled, and
size-independent properties at the BASELINE.json sizes.

Bars (BASELINE.json north_star): symbolic structure bit-exact; factor entries within 1e-10 relative (0/0 := 0, explicit
zeros must stay exactly zero); backward error ||Ax-b||/(||A|| ||x||) <= 1e-12."""
import os

import numpy as np
import pytest

import sforacle
import sfref
import sparseframe_b200 as sf
from conftest import build_matrix

pytestmark = pytest.mark.gpu

ENTRY_TOL = 1e-10
BACKWARD_TOL = 1e-12
ENV_KEYS = ("SFB200_NB", "SFB200_OUTER", "SFB200_OUTER2", "SFB200_SMALL_A", "SFB200_SMALL_B", "SFB200_GEMM128_MIN",
            "SFB200_WS_KMAJOR", "SFB200_LU_UT_COPY")


@pytest.fixture(autouse=True)
def _clean_env():
    for k in ENV_KEYS:
        os.environ.pop(k, None)
    yield
    for k in ENV_KEYS:
        os.environ.pop(k, None)


def gpu_factor(S, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = str(v)
    plan = sf.Plan(S)
    split = S.kind == 1 and S.int("symmetric") == 0
    lsx = plan.factorize(S.array("Lx"), S.array("Ux") if split else None)
    return plan, lsx


def oracle_structure(S):
    st = {k: S.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
    if S.kind == 1 and S.int("symmetric") == 0:
        st["Up"], st["Ui"] = S.array("Up"), S.array("Ui")
    return st


def compare(kind, st, ref_lsx, lsx, tol=ENTRY_TOL):
    mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], len(ref_lsx))
    a, b = ref_lsx[mask], lsx[mask]
    assert not np.isnan(b).any()
    assert sf.entrywise_rel_diff(a, b) <= tol
    assert np.array_equal(a == 0, b == 0), "explicit zeros must stay exactly zero"


ENVS = [dict(), dict(SFB200_SMALL_A=0, SFB200_SMALL_B=0), dict(SFB200_NB=16, SFB200_OUTER=32, SFB200_SMALL_A=0, SFB200_SMALL_B=0),
        dict(SFB200_NB=8, SFB200_OUTER=16, SFB200_SMALL_B=3000), dict(SFB200_GEMM128_MIN=32)]


@pytest.mark.parametrize("env", ENVS, ids=[str(e) for e in ENVS])
def test_golden_vectors(golden, env):
    A, _ = build_matrix(golden["gen"], golden["grid"])
    S = sf.analyze(A, golden["kind"], devslot=golden["devslot"])
    for k in ("Perm", "Super", "SuperMap", "Lsip", "Lsxp", "Lsi"):
        assert np.array_equal(S.array(k), golden[k]), k
    plan, lsx = gpu_factor(S, env)
    compare(golden["kind"], golden, golden["Lsx"], lsx)
    res, backward = sforacle.validate(golden["kind"], oracle_structure(S), S.array("Lx"), lsx,
                                      S.array("Ux") if "Ux" in golden else None)
    assert backward <= BACKWARD_TOL
    plan.close()


CASES = [("lap2d", 120, "cholesky"), ("lap3d", 24, "cholesky"), ("el3d", 8, "cholesky"), ("cd3d", 20, "lu"),
         ("lap3d", 16, "lu"), ("lap2d", 64, "lu")]


@pytest.mark.parametrize("gen,g,kind", CASES)
@pytest.mark.parametrize("env", [dict(), dict(SFB200_NB=32, SFB200_OUTER=64, SFB200_SMALL_B=8192)], ids=["default", "nb32"])
def test_against_oracle_restatement(gen, g, kind, env):
    A, _ = build_matrix(gen, g)
    S = sf.analyze(A, kind)
    st = oracle_structure(S)
    ux = S.array("Ux") if "Up" in st else None
    ref_lsx, bad = sforacle.factorize(kind, st, S.array("Lx"), ux)
    assert bad == 0
    plan, lsx = gpu_factor(S, env)
    compare(kind, st, ref_lsx, lsx)
    _, backward = sforacle.validate(kind, st, S.array("Lx"), lsx, ux)
    assert backward <= BACKWARD_TOL
    plan.close()


@pytest.mark.skipif(not (sfref.available("cholesky") and sfref.available("lu")), reason="compiled reference did not travel")
@pytest.mark.parametrize("gen,g,kind", [("lap3d", 40, "cholesky"), ("lap2d", 300, "cholesky"), ("cd3d", 32, "lu"),
                                        ("el3d", 12, "cholesky")])
def test_against_compiled_reference(gen, g, kind):
    A, (n, r, c, x, sym) = build_matrix(gen, g)
    ref = sfref.Reference(n, r, c, x, sym, kind=kind).analyze().factorize()
    S = sf.analyze(A, kind)
    for k in ("Perm", "Super", "Lsip", "Lsxp", "Lsi"):
        assert np.array_equal(S.array(k), ref.array(k)[:len(S.array(k))]), k
    plan, lsx = gpu_factor(S)
    compare(kind, oracle_structure(S), ref.array("Lsx"), lsx)
    ref.close()
    plan.close()


def test_foreign_structure_drop_in(golden):
    """The drop-in boundary: the structure comes from somebody else's analysis (here: the golden arrays dumped from the
    reference's SparseFrame_analyze), only the numeric step is ours."""
    split = "Up" in golden
    W = sf.api.wrap_structure(golden["kind"], golden["n"], golden["Super"], golden["Lsip"], golden["Lsxp"],
                              golden["Lsi"], golden["Lp"], golden["Li"], golden["Up"] if split else None,
                              golden["Ui"] if split else None)
    plan = sf.Plan(W)
    lsx = plan.factorize(golden["Lx"], golden["Ux"] if split else None)
    compare(golden["kind"], golden, golden["Lsx"], lsx)
    plan.close()


def test_edge_cases():
    # 1x1
    S = sf.analyze(sf.compress(1, [0], [0], [4.0], True), "cholesky")
    plan, lsx = gpu_factor(S)
    assert lsx.tolist() == [2.0]
    plan.close()
    # diagonal matrix (no off-diagonals: METIS skipped, every column its own chain)
    n = 9
    d = np.arange(1.0, n + 1)
    S = sf.analyze(sf.compress(n, np.arange(n), np.arange(n), d, True), "cholesky")
    plan, lsx = gpu_factor(S)
    st = oracle_structure(S)
    ref_lsx, _ = sforacle.factorize("cholesky", st, S.array("Lx"))
    compare("cholesky", st, ref_lsx, lsx)
    plan.close()
    # ragged: an arrow matrix (one dense row/column) gives one wide last supernode and many 1-column ones
    n = 40
    rows = np.concatenate([np.arange(n), np.full(n - 1, n - 1)])
    cols = np.concatenate([np.arange(n), np.arange(n - 1)])
    vals = np.concatenate([np.full(n, 50.0), np.full(n - 1, -1.0)])
    S = sf.analyze(sf.compress(n, rows, cols, vals, True), "cholesky")
    plan, lsx = gpu_factor(S)
    st = oracle_structure(S)
    ref_lsx, _ = sforacle.factorize("cholesky", st, S.array("Lx"))
    compare("cholesky", st, ref_lsx, lsx)
    plan.close()


@pytest.mark.parametrize("gen,g,kind", [("lap2d", 40, "cholesky"), ("cd3d", 10, "lu"), ("lap3d", 9, "lu")])
def test_duplicate_entries_last_one_wins(gen, g, kind):
    """Several copies of one (i, j): the reference's loadA assigns in storage order, the last copy stays (CH.c:2018,
    LU.c:2498/2513).  k_assemble scatters all entries concurrently, so the plan must have switched the earlier copies off;
    every 3rd entry gets a wrong-valued earlier twin here and the factor must equal the oracle's, run after run."""
    n, r, c, x, sym = sf.matrices.generate(gen, g)
    twin = np.arange(0, len(r), 3)
    r2, c2 = np.concatenate([r[twin], r]), np.concatenate([c[twin], c])
    x2 = np.concatenate([x[twin] * 3.0 + 1.0, x])
    S = sf.analyze(sf.compress(n, r2, c2, x2, sym), kind)
    st = oracle_structure(S)
    ux = S.array("Ux") if "Up" in st else None
    ref_lsx, bad = sforacle.factorize(kind, st, S.array("Lx"), ux)
    assert bad == 0
    plan = sf.Plan(S)
    for _ in range(3):
        compare(kind, st, ref_lsx, plan.factorize(S.array("Lx"), ux))
    plan.close()


def test_indefinite_matrix_is_reported():
    n, r, c, x, sym = sf.matrices.generate("lap2d", 10)
    x = x.copy()
    x[r == c] = -4.0
    S = sf.analyze(sf.compress(n, r, c, x, sym), "cholesky")
    plan = sf.Plan(S)
    with pytest.raises(RuntimeError, match="pivot"):
        plan.factorize(S.array("Lx"))
    plan.close()


def test_device_resident_api_and_repeatability():
    import torch
    A, _ = build_matrix("lap3d", 20)
    S = sf.analyze(A, "cholesky")
    plan = sf.Plan(S)
    host = plan.factorize(S.array("Lx"))
    d_lx = torch.from_numpy(S.array("Lx")).cuda()
    plan.set_stream(torch.cuda.current_stream().cuda_stream)
    plan.factorize_device(d_lx.data_ptr())
    plan.sync()
    dev = plan.download()
    st = oracle_structure(S)
    # atomics make the summation order vary run to run: repeatable to rounding, not bitwise
    compare("cholesky", st, host, dev, tol=1e-12)
    assert plan.last_launches() > 0 and plan.last_ms(0) > 0
    prof = plan.profile()
    assert sum(v["launches"] for v in prof.values()) == plan.int("nlaunches")
    plan.close()


def test_linearity_scaling_property():
    """chol(4A) = 2 chol(A): scaling by a power of 4 is exact in binary floating point, so the two factors must agree
    to summation-order rounding."""
    A, _ = build_matrix("lap3d", 24)
    S = sf.analyze(A, "cholesky")
    plan = sf.Plan(S)
    l1 = plan.factorize(S.array("Lx"))
    l4 = plan.factorize(4.0 * S.array("Lx"))
    st = oracle_structure(S)
    compare("cholesky", st, 2.0 * l1, l4, tol=1e-12)
    plan.close()


@pytest.mark.parametrize("cfg", ["cfg1", "cfg2"])
def test_full_size_backward_error(cfg):
    """BASELINE.json configs[0] / configs[1] at full size: the oracle would take minutes, so check the
    size-independent property -- solve with the factor and measure the backward error of the permuted system."""
    gen, g, kind = sf.matrices.CONFIGS[cfg]
    A, _ = build_matrix(gen, g)
    S = sf.analyze(A, kind)
    plan = sf.Plan(S)
    lsx = plan.factorize(S.array("Lx"))
    res, backward = sforacle.validate(kind, oracle_structure(S), S.array("Lx"), lsx)
    assert backward <= BACKWARD_TOL
    assert res <= 1e-13
    plan.close()


@pytest.mark.parametrize("gen,g,kind", [("lap3d", 30, "cholesky"), ("cd3d", 24, "lu")])
def test_graph_replays_match_the_first_plain_run(gen, g, kind):
    """The engine issues the first factorization of a plan plainly, captures the second one as a CUDA graph and replays
    that graph afterwards (device-resident path, streamed copy-out into one page-locked buffer).  Every replay, also
    with new matrix values, must give what a fresh plan gives."""
    import torch
    A, _ = build_matrix(gen, g)
    S = sf.analyze(A, kind)
    st = oracle_structure(S)
    ux = S.array("Ux") if "Up" in st else None
    lx = S.array("Lx")
    ref_lsx, bad = sforacle.factorize(kind, st, lx, ux)
    assert bad == 0
    plan = sf.Plan(S)
    pinned = torch.empty(S.int("xsize"), dtype=torch.float64).pin_memory()
    for it in range(4):                                    # plain, capture, replay, replay
        pinned.zero_()
        out = plan.factorize(lx, ux, out=pinned.numpy())   # streamed copy-out (page-locked destination)
        compare(kind, st, ref_lsx, out)
    for it in range(4):                                    # pageable destination: device path + pack + one copy
        compare(kind, st, ref_lsx, plan.factorize(lx, ux))
    # new values through the captured graphs: 3 A -> sqrt(3) L (Cholesky), L unchanged and 3 U (LU)
    scaled = plan.factorize(3.0 * lx, None if ux is None else 3.0 * ux)
    fresh = sf.Plan(S)
    compare(kind, st, fresh.factorize(3.0 * lx, None if ux is None else 3.0 * ux), scaled, tol=1e-12)
    fresh.close()
    plan.close()


@pytest.mark.parametrize("cfg", ["cfg3", "cfg4", "cfg5"])
def test_full_size_device_residual(cfg):
    """BASELINE.json configs[2], [3], [4] at full size on ONE GPU -- 27-point 3-dof elasticity 80^3 (36.9 GB factor),
    LU of the 3-D convection-diffusion 100^3 (14 GB factor, fronts wide enough for every blocking level and for the
    look-ahead split) and the 160^3 Laplacian (58.4 GB factor, the multi-GPU config, which fits one B200): factorize
    without copy-out, solve and take the reference's residual (CH.c:3197-3263) on the device (the device solve itself is
    checked against the oracle in test_device_solve_and_residual)."""
    gen, g, kind = sf.matrices.CONFIGS[cfg]
    A, _ = build_matrix(gen, g)
    S = sf.analyze(A, kind)
    plan = sf.Plan(S)
    split = kind == "lu" and not A.symmetric
    plan.factorize(S.array("Lx"), S.array("Ux") if split else None, download=False)
    b = 1.0 + np.arange(A.n) / A.n
    x = plan.solve(b)
    dev = plan.residual(x, b)
    assert np.isfinite(x).all()
    assert dev["rnorm"] / (dev["anorm"] * dev["xnorm"]) <= BACKWARD_TOL
    assert dev["residual"] <= 1e-13
    plan.close()


@pytest.mark.skipif(not sfref.available("cholesky"), reason="compiled reference did not travel")
@pytest.mark.skipif(os.environ.get("SFB200_SKIP_SLOW") == "1", reason="SFB200_SKIP_SLOW=1")
def test_cfg2_entrywise_against_compiled_reference():
    """The BASELINE config itself (3-D Laplacian 100^3, 1.1e9 factor entries) entry by entry against the compiled
    reference's CPU factorization (about a minute of host time): structure bit-exact, every entry within 1e-10 relative,
    explicit zeros exactly zero.  Compared in slices so that the temporaries stay small next to the two 8.8 GB factors."""
    gen, g, kind = sf.matrices.CONFIGS["cfg2"]
    A, (n, r, c, x, sym) = build_matrix(gen, g)
    ref = sfref.Reference(n, r, c, x, sym, kind=kind, blas_threads=min(os.cpu_count(), 16)).analyze().factorize()
    S = sf.analyze(A, kind)
    for k in ("Perm", "Super", "Lsip", "Lsxp", "Lsi"):
        assert np.array_equal(S.array(k), ref.array(k)[:len(S.array(k))]), k
    ref_lsx = ref.array("Lsx")
    ref.close()
    plan, lsx = gpu_factor(S)
    plan.close()
    st = oracle_structure(S)
    mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], len(ref_lsx))
    worst = 0.0
    step = 1 << 26
    for i in range(0, len(ref_lsx), step):
        m = mask[i:i + step]
        a, b = ref_lsx[i:i + step][m], lsx[i:i + step][m]
        assert not np.isnan(b).any()
        assert np.array_equal(a == 0, b == 0), "explicit zeros must stay exactly zero"
        worst = max(worst, sf.entrywise_rel_diff(a, b))
    print("cfg2 entrywise max relative difference vs the compiled reference: %.3e over %d entries" % (worst, int(mask.sum())))
    assert worst <= ENTRY_TOL


@pytest.mark.parametrize("gen,g,kind", [("cd3d", 30, "lu"), ("lap3d", 30, "lu"), ("lap3d", 34, "cholesky")])
def test_lookahead_split_with_narrow_blocking(gen, g, kind):
    """Top fronts of ~900-1150 columns with a 64/128/256 blocking: four or more top-level block columns, so the trailing
    updates are split between the main and the side stream (look-ahead), for LU as well as Cholesky."""
    env = dict(SFB200_OUTER="128", SFB200_OUTER2="256")
    os.environ.update(env)
    try:
        A, _ = build_matrix(gen, g)
        S = sf.analyze(A, kind)
        st = oracle_structure(S)
        ux = S.array("Ux") if "Up" in st else None
        plan = sf.Plan(S)
    finally:
        for k in env:
            os.environ.pop(k, None)
    ref_lsx, bad = sforacle.factorize(kind, st, S.array("Lx"), ux)
    assert bad == 0
    for _ in range(3):                     # a race would not need to show on the first run
        compare(kind, st, ref_lsx, plan.factorize(S.array("Lx"), ux))
    plan.close()


@pytest.mark.parametrize("gen,g,kind", [("cd3d", 30, "lu"), ("lap3d", 30, "lu")])
@pytest.mark.parametrize("env", [dict(), dict(SFB200_OUTER=128, SFB200_OUTER2=256), dict(SFB200_LU_UT_COPY=0, SFB200_WS_KMAJOR=1),
                                 dict(SFB200_LU_UT_COPY=0)], ids=["ut-copy", "ut-copy-narrow", "kmajor-ws", "kmajor-classic"])
def test_lu_in_panel_updates_operand_paths(gen, g, kind, env):
    """LU's in-panel updates C -= L21 * U12 need rows of U11, which the panel stores k-major.  Default: the row-block solves
    keep a transposed copy (U11^T) in a scratch area and the updates read it row-major on the warp-specialized kernel.
    SFB200_LU_UT_COPY=0: k-major operand on the classic cp.async kernel, or (SFB200_WS_KMAJOR=1) staged column by column
    with 128-byte bulk copies on the warp-specialized one.  Same factor on all three paths, fronts of 900-1450 columns."""
    A, _ = build_matrix(gen, g)
    S = sf.analyze(A, kind)
    st = oracle_structure(S)
    ux = S.array("Ux") if "Up" in st else None
    ref_lsx, bad = sforacle.factorize(kind, st, S.array("Lx"), ux)
    assert bad == 0
    plan, lsx = gpu_factor(S, env)
    compare(kind, st, ref_lsx, lsx)
    for _ in range(2):
        compare(kind, st, ref_lsx, plan.factorize(S.array("Lx"), ux))
    _, backward = sforacle.validate(kind, st, S.array("Lx"), lsx, ux)
    assert backward <= BACKWARD_TOL
    plan.close()


def test_streamed_copy_out_into_pinned_memory():
    """Page-locked destination: the factor is copied out level by level (and block column by block column for the wide
    fronts) behind the factorization.  Must give the same factor as the pack-then-copy path used for pageable memory."""
    import torch
    for gen, g, kind, env in (("lap3d", 28, "cholesky", dict(SFB200_OUTER=128, SFB200_OUTER2=256)),
                              ("lap2d", 200, "cholesky", dict()), ("cd3d", 22, "lu", dict(SFB200_OUTER=128, SFB200_OUTER2=0))):
        for k, v in env.items():
            os.environ[k] = str(v)
        A, _ = build_matrix(gen, g)
        S = sf.analyze(A, kind)
        split = kind == "lu" and S.int("symmetric") == 0
        plan = sf.Plan(S)
        ux = S.array("Ux") if split else None
        pageable = plan.factorize(S.array("Lx"), ux)
        pinned = torch.full((S.int("xsize"),), float("nan"), dtype=torch.float64).pin_memory()
        out = plan.factorize(S.array("Lx"), ux, out=pinned.numpy())
        st = oracle_structure(S)
        assert not np.isnan(out).any()
        compare(kind, st, pageable, out, tol=1e-12)
        ref_lsx, bad = sforacle.factorize(kind, st, S.array("Lx"), ux)
        compare(kind, st, ref_lsx, out)
        plan.close()
        for k in env:
            os.environ.pop(k, None)


@pytest.mark.parametrize("gen,g,kind", [("lap3d", 24, "cholesky"), ("lap2d", 150, "cholesky"), ("cd3d", 20, "lu"),
                                        ("lap3d", 14, "lu"), ("el3d", 8, "cholesky")])
def test_device_solve_and_residual(gen, g, kind):
    """Row f1: forward/backward substitution and the reference's residual check on the GPU, against the oracle's
    restatement of SparseFrame_solve_supernodal / SparseFrame_validate."""
    A, _ = build_matrix(gen, g)
    S = sf.analyze(A, kind)
    st = oracle_structure(S)
    split = "Up" in st
    ux = S.array("Ux") if split else None
    plan = sf.Plan(S)
    lsx = plan.factorize(S.array("Lx"), ux)
    n = A.n
    b = 1.0 + np.arange(n) / n                      # the reference's right-hand side (CH.c:3186)
    x = plan.solve(b)
    x_ref = sforacle.solve(kind, st, lsx, b)
    assert np.max(np.abs(x - x_ref)) <= 1e-11 * np.max(np.abs(x_ref))
    dev = plan.residual(x, b)
    res_ref, norms = sforacle.residual(kind, st, S.array("Lx"), b, x, ux)
    assert dev["anorm"] == pytest.approx(norms[0], rel=1e-13)
    assert dev["bnorm"] == norms[1] and dev["xnorm"] == norms[2]
    assert dev["rnorm"] <= 1e-12 * dev["anorm"] * dev["xnorm"]
    assert dev["rnorm"] / (dev["anorm"] * dev["xnorm"]) <= BACKWARD_TOL
    # several right-hand sides at once
    B = np.asfortranarray(np.stack([b, np.cos(np.arange(n)), np.ones(n)], axis=1))
    Xm = plan.solve(B)
    for k in range(3):
        xk = sforacle.solve(kind, st, lsx, B[:, k])
        assert np.max(np.abs(Xm[:, k] - xk)) <= 1e-11 * np.max(np.abs(xk))
    assert plan.last_ms(2) > 0
    plan.close()
