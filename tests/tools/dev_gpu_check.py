"""Developer GPU check: FP64 pipe peaks, parity of the CUDA path against the compiled reference on small cases under
several plan parameterisations, then one timed mid-size and (optionally) full-size factorization."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sparseframe_b200 as sf  # noqa: E402
import sfref  # noqa: E402


def run_case(kind, g, fk, env=None, check_ref=True, repeat=1, backward=False):
    env = env or {}
    for k in ("SFB200_NB", "SFB200_OUTER", "SFB200_SMALL_A", "SFB200_SMALL_B", "SFB200_GEMM128_MIN"):
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    n, r, c, x, sym = sf.matrices.generate(kind, g)
    A = sf.compress(n, r, c, x, sym)
    t0 = time.time()
    S = sf.analyze(A, fk)
    ta = time.time() - t0
    t0 = time.time()
    plan = sf.Plan(S)
    tp = time.time() - t0
    split = fk == "lu" and not sym
    lx, ux = S.array("Lx"), (S.array("Ux") if split else None)
    msg = "%s g=%d %s env=%s n=%d nsuper=%d levels=%d launches=%d small=%d large=%d | analyze %.2fs plan %.2fs" % (
        kind, g, fk, env, n, S.int("nsuper"), plan.int("nlevels"), plan.int("nlaunches"), plan.int("nsmall"),
        plan.int("nlarge"), ta, tp)
    lsx = None
    for _ in range(repeat):
        lsx = plan.factorize(lx, ux)
    msg += " | numeric %.3f ms e2e %.3f ms  %.1f GFLOP/s" % (plan.last_ms(0), plan.last_ms(1),
                                                            S.flops / (plan.last_ms(0) * 1e-3) / 1e9)
    if check_ref:
        ref = sfref.Reference(n, r, c, x, sym, kind=fk).analyze().factorize()
        st = {k: ref.array(k) for k in ("Super", "Lsip", "Lsxp", "Lsi")}
        for k in st:
            assert np.array_equal(st[k], S.array(k)[:len(st[k])]), k
        mask = sf.lower_trapezoid_mask(fk, st["Super"], st["Lsip"], st["Lsxp"], ref.long("xsize"))
        a, b = ref.array("Lsx")[mask], lsx[mask]
        rel = sf.entrywise_rel_diff(a, b)
        zm = int(((a == 0) != (b == 0)).sum())
        msg += " | max rel diff %.2e zero-mismatch %d nan %d" % (rel, zm, int(np.isnan(b).sum()))
        ref.close()
    if backward:
        res = sf.validate(fk, n, S.array("Lp"), S.array("Li"), lx, S.array("Super"), S.array("Lsip"), S.array("Lsxp"),
                          S.array("Lsi"), lsx, *((S.array("Up"), S.array("Ui"), ux) if split else ()))
        msg += " | residual %.2e backward %.2e" % res
    print(msg, flush=True)
    plan.close()
    S.close()


def main():
    print("devices", sf.device_count(), flush=True)
    for which in ("dfma", "dmma"):
        print("fp64 peak %s: %.2f TFLOP/s" % (which, sf.fp64_peak(0, which, 8192)), flush=True)
    big = len(sys.argv) > 1 and sys.argv[1] == "big"
    cases = [
        ("lap2d", 20, "cholesky", {}),
        ("lap3d", 12, "cholesky", {}),
        ("lap3d", 12, "cholesky", dict(SFB200_SMALL_A=0, SFB200_SMALL_B=0)),
        ("lap3d", 16, "cholesky", dict(SFB200_NB=16, SFB200_OUTER=32, SFB200_SMALL_A=0, SFB200_SMALL_B=0)),
        ("lap2d", 100, "cholesky", dict(SFB200_NB=8, SFB200_OUTER=16, SFB200_SMALL_B=2000)),
        ("el3d", 6, "cholesky", dict(SFB200_NB=32, SFB200_OUTER=64)),
        ("lap3d", 24, "cholesky", {}),
        ("lap3d", 24, "cholesky", dict(SFB200_GEMM128_MIN=64)),
        ("cd3d", 12, "lu", {}),
        ("cd3d", 12, "lu", dict(SFB200_SMALL_A=0, SFB200_SMALL_B=0)),
        ("cd3d", 16, "lu", dict(SFB200_NB=16, SFB200_OUTER=32, SFB200_SMALL_A=0, SFB200_SMALL_B=0)),
        ("lap3d", 12, "lu", dict(SFB200_NB=8, SFB200_OUTER=16, SFB200_SMALL_B=3000)),
        ("cd3d", 20, "lu", dict(SFB200_GEMM128_MIN=64)),
    ]
    for kind, g, fk, env in cases:
        try:
            run_case(kind, g, fk, env)
        except Exception as e:  # noqa: BLE001
            print("FAILED", kind, g, fk, env, repr(e), flush=True)
    for kind, g, fk in (("lap3d", 40, "cholesky"), ("lap2d", 300, "cholesky"), ("cd3d", 32, "lu")):
        try:
            run_case(kind, g, fk, {}, check_ref=True, repeat=3, backward=True)
        except Exception as e:  # noqa: BLE001
            print("FAILED", kind, g, fk, repr(e), flush=True)
    if big:
        for kind, g, fk in (("lap3d", 100, "cholesky"), ("lap2d", 1000, "cholesky")):
            try:
                run_case(kind, g, fk, {}, check_ref=False, repeat=3, backward=True)
            except Exception as e:  # noqa: BLE001
                print("FAILED", kind, g, fk, repr(e), flush=True)


if __name__ == "__main__":
    main()
