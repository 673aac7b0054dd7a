"""Differential run of the engine's host side against the compiled reference on random irregular matrices (no GPU):
symbolic structure bit for bit, the CPU-interpreted schedule's factor entry by entry.  Real and complex, Cholesky and LU,
random slot sizes (capped supernodes, many stages), arrows / bands / forests / duplicates / shuffled file order, symmetric
files with entries in the upper triangle, missing diagonal entries (analysis only).
    python tests/tools/differential_fuzz.py [cases] [seed]
Every case is printed before it runs, so that a crash inside the reference names its input."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sforacle  # noqa: E402
import sfref  # noqa: E402
import sparseframe_b200 as sf  # noqa: E402

INT_ARRAYS = ("Perm", "Parent", "Post", "ColCount", "Lp", "Li", "LTp", "LTi", "Super", "SuperMap", "Lsip", "Lsxp", "Lsi",
              "ST_Map", "ST_Pointer", "ST_Index", "Aoffset", "Moffset")


def pattern(rng, n, shape):
    if shape == "random":
        m = int(rng.uniform(0.5, 4.0) * n)
        i, j = rng.integers(0, n, m), rng.integers(0, n, m)
    elif shape == "band":
        w = int(rng.integers(1, 6))
        i = np.repeat(np.arange(n), w)
        j = i - np.tile(np.arange(1, w + 1), n)
    elif shape == "arrow":
        k = int(rng.integers(1, 4))
        i = np.concatenate([np.repeat(np.arange(n - k, n), n), np.arange(1, n)])
        j = np.concatenate([np.tile(np.arange(n), k), np.arange(0, n - 1)])
    else:                                   # forest: blocks that do not touch
        b = int(rng.integers(2, 6))
        m = int(2.0 * n)
        i, j = rng.integers(0, n, m), rng.integers(0, n, m)
        keep = (i * b // n) == (j * b // n)
        i, j = i[keep], j[keep]
    keep = (i > j) & (j >= 0)
    i, j = i[keep], j[keep]
    key = np.unique(i * n + j)
    return key // n, key % n


def make(rng, n, shape, symmetric, cplx, dup):
    i, j = pattern(rng, n, shape)
    v = -rng.random(len(i)) - 0.1
    if symmetric:
        rows, cols, vals = i, j, v
    else:
        up = rng.random(len(i)) < rng.choice([1.0, 0.7, 0.3])
        rows = np.concatenate([i, j[up]])
        cols = np.concatenate([j, i[up]])
        vals = np.concatenate([v, -rng.random(int(up.sum())) - 0.1])
    rs, cs = np.zeros(n), np.zeros(n)
    np.add.at(rs, rows, np.abs(vals))
    np.add.at(cs, cols, np.abs(vals))
    if symmetric:
        dom = rs + cs
    else:
        dom = np.maximum(rs, cs)
    rows = np.concatenate([rows, np.arange(n)])
    cols = np.concatenate([cols, np.arange(n)])
    vals = np.concatenate([vals, dom + 1.0])
    if cplx:
        ph = rng.uniform(-0.7, 0.7, len(vals))
        if symmetric == "hermitian":
            ph[rows == cols] = 0.0
        vals = vals * np.exp(1j * ph)
    order = rng.permutation(len(rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    if dup:                                 # earlier copies of some entries with other values: the last copy wins
        pick = rng.choice(len(rows), max(1, len(rows) // 10), replace=False)
        rows = np.concatenate([rows[pick], rows])
        cols = np.concatenate([cols[pick], cols])
        vals = np.concatenate([vals[pick] * 0.5, vals])
    return rows.astype(np.int64), cols.astype(np.int64), vals


def one_case(rng, k):
    kind = str(rng.choice(["cholesky", "lu"]))
    cplx = bool(rng.random() < 0.5)
    symmetric = True if kind == "cholesky" else bool(rng.random() < 0.3)
    n = int(rng.choice([1, 2, 3, 5, 17, 60, 150, 400, 900]))
    shape = str(rng.choice(["random", "band", "arrow", "forest"]))
    dup = bool(rng.random() < 0.3)
    slot = int(rng.choice([sf.DEFAULT_DEVSLOT, 200_000, 20_000, 3_000]))
    if n <= 3:
        shape = "band"                      # the reference's METIS call on a 2- or 3-vertex graph is not a permutation
    print("case %d: %s %s n=%d %s sym=%s dup=%s slot=%d" % (k, kind, "complex" if cplx else "real", n, shape, symmetric, dup, slot), flush=True)
    herm = "hermitian" if (kind == "cholesky") else symmetric
    r, c, v = make(rng, n, shape, herm if cplx else symmetric, cplx, dup)
    if n in (2, 3) and len(r) > n:
        return "skipped"                    # see above
    structure_only = False
    if symmetric and n > 3 and rng.random() < 0.2:       # a "symmetric" file listing some entries in the upper triangle
        flip = (r != c) & (rng.random(len(r)) < 0.4)
        r, c = np.where(flip, c, r), np.where(flip, r, c)
    if n > 3 and not dup and rng.random() < 0.15:          # missing diagonal entries: only the analysis can be compared
        drop = (r == c) & (rng.random(len(r)) < 0.3)
        r, c, v = r[~drop], c[~drop], v[~drop]
        structure_only = True
    if os.environ.get("SFFUZZ_DUMP"):       # keep the input of the case about to run (a crash inside the reference kills us)
        np.savez(os.environ["SFFUZZ_DUMP"], n=n, r=r, c=c, v=v, kind=kind, symmetric=symmetric, slot=slot, cplx=cplx)
    A = sf.compress(n, r, c, v, symmetric)
    S = sf.analyze(A, kind, devslot=slot)
    if dup:
        # duplicate entries: the reference's METIS input has repeated neighbours (it crashes, hangs or mis-orders on some of
        # them); the engine removes them, so its ordering is its own.  Checked against the oracle's restatement on the
        # engine's own structure instead: the last copy of every entry must be the one that counts.
        assert np.array_equal(np.sort(S.array("Perm")), np.arange(n))
        st = {q: S.array(q) for q in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
        st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
        split = kind == "lu" and not symmetric
        if split:
            st["Up"], st["Ui"] = S.array("Up"), S.array("Ui")
        fact = sforacle.zfactorize if cplx else sforacle.factorize
        ref, bad = fact(kind, st, S.array("Lx"), S.array("Ux") if split else None)
        assert bad == 0
        if cplx:
            out, bad, _, _ = sforacle.check_plan_complex(kind, n, A.Cp, A.Ci, A.Cx, symmetric, slot, row_phase=bool(rng.random() < 0.5))
        else:
            out, bad, _ = sforacle.plan_check(kind, n, A.Cp, A.Ci, A.Cx, symmetric, slot)
        mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], S.int("xsize"))
        scale = np.abs(ref).max() if ref.size else 1.0
        assert bad == 0 and (ref.size == 0 or np.abs(out[mask] - ref[mask]).max() <= 1e-11 * scale)
        return "dup"
    if S.array("ColCount").max() == n:
        # a column whose count is n: the reference's weighted postorder writes its bucket array out of bounds
        # (CH.c:1177-1200) and corrupts its heap -- nothing to compare with
        return "skipped"
    R = sfref.Reference(n, r, c, v, symmetric, kind=kind, devslot=slot, blas_threads=1).analyze()
    names = list(INT_ARRAYS) + (["Up", "Ui", "UTp", "UTi"] if (kind == "lu" and not symmetric) else [])
    for name in names:
        a, b = S.array(name), R.array(name)
        assert np.array_equal(a, b[:len(a)]), (name, a[:10], b[:10])
    for name in ("nsuper", "isize", "xsize", "csize", "nstage", "nsleaf"):
        assert S.int(name) == R.long(name), name
    assert np.array_equal(S.array("LeafQueue")[:S.int("nsleaf")], R.array("LeafQueue")[:S.int("nsleaf")])
    assert np.array_equal(S.array("Lx"), R.array("Lx"))
    if structure_only:
        R.close()
        return "ok"
    R.factorize()
    ref = R.array("Lsx")
    mask = sf.lower_trapezoid_mask(kind, R.array("Super"), R.array("Lsip"), R.array("Lsxp"), R.long("xsize"))
    prm = dict(nb=8, outer=16, small_a=0, small_b=0) if rng.random() < 0.5 else {}
    if cplx:
        out, bad, _, _ = sforacle.check_plan_complex(kind, n, A.Cp, A.Ci, A.Cx, symmetric, slot, row_phase=bool(rng.random() < 0.5), **prm)
    else:
        out, bad, _ = sforacle.plan_check(kind, n, A.Cp, A.Ci, A.Cx, symmetric, slot, **prm)
    assert bad == 0
    scale = np.abs(ref).max() if ref.size else 1.0
    assert ref.size == out.size and (ref.size == 0 or np.abs(out[mask] - ref[mask]).max() <= 1e-11 * scale), "factor"
    assert np.array_equal(out[mask] == 0, ref[mask] == 0), "zero pattern"
    R.close()
    return "ok"


if __name__ == "__main__":
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    done = {"ok": 0, "skipped": 0, "dup": 0}
    for k in range(cases):
        done[one_case(rng, k)] += 1
    print("differential fuzz: %d cases identical to the compiled reference, %d with duplicate entries equal to the oracle on "
          "the engine's own structure, %d skipped" % (done["ok"], done["dup"], done["skipped"]))
