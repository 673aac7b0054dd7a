"""Deterministic synthetic PDE matrices for the five BASELINE.json configs.

No RNG anywhere: every entry is a closed-form function of the grid position, so the same call
gives the same triplets on every box.  Vertex id ``v = (i*g + j)*g + k`` (2-D: ``v = i*g + j``).

Each generator returns ``(n, rows, cols, vals, symmetric)`` as 0-based int64 / float64 triplets in
the order a MatrixMarket file would list them.  Symmetric matrices carry the lower triangle only
(row >= col), which is what the reference's reader expects (``Cholesky/Source/SparseFrame.c:400-524``:
entries are taken as given, duplicates are not summed).
"""
from __future__ import annotations

import itertools

import numpy as np


def laplace2d(g: int):
    """cfg1: 2-D 5-point Laplacian, Dirichlet truncation; diag 4, off-diag -1."""
    n = g * g
    v = np.arange(n, dtype=np.int64).reshape(g, g)
    r = [v.ravel(), v[:, 1:].ravel(), v[1:, :].ravel()]
    c = [v.ravel(), v[:, :-1].ravel(), v[:-1, :].ravel()]
    x = [np.full(n, 4.0), np.full(g * (g - 1), -1.0), np.full(g * (g - 1), -1.0)]
    return n, np.concatenate(r), np.concatenate(c), np.concatenate(x), True


def _axis_pairs(v, a):
    hi = [slice(None)] * 3
    lo = [slice(None)] * 3
    hi[a] = slice(1, None)
    lo[a] = slice(None, -1)
    return v[tuple(hi)].ravel(), v[tuple(lo)].ravel()


def laplace3d(g: int):
    """cfg2 / cfg5: 3-D 7-point Laplacian; diag 6, off-diag -1."""
    n = g ** 3
    v = np.arange(n, dtype=np.int64).reshape(g, g, g)
    r, c, x = [v.ravel()], [v.ravel()], [np.full(n, 6.0)]
    for a in range(3):
        hi, lo = _axis_pairs(v, a)
        r.append(hi)
        c.append(lo)
        x.append(np.full(hi.size, -1.0))
    return n, np.concatenate(r), np.concatenate(c), np.concatenate(x), True


def convdiff3d(g: int, conv=(0.3, 0.2, 0.1)):
    """cfg4: unsymmetric 3-D 7-point convection-diffusion; diag 6, (v+stride_a, v) = -1-c_a,
    (v, v+stride_a) = -1+c_a.  Irreducibly diagonally dominant => safe without pivoting."""
    n = g ** 3
    v = np.arange(n, dtype=np.int64).reshape(g, g, g)
    r, c, x = [v.ravel()], [v.ravel()], [np.full(n, 6.0)]
    for a, ca in zip(range(3), conv):
        hi, lo = _axis_pairs(v, a)
        r += [hi, lo]
        c += [lo, hi]
        x += [np.full(hi.size, -1.0 - ca), np.full(hi.size, -1.0 + ca)]
    return n, np.concatenate(r), np.concatenate(c), np.concatenate(x), False


def elasticity3d(g: int):
    """cfg3: 27-point, 3-dof-per-node elasticity-like SPD matrix; dof id = 3*v + c.

    Node pairs within the 3x3x3 neighbourhood are coupled by ``-w(d) * B`` with
    ``B = [[1,.1,.1],[.1,1,.1],[.1,.1,1]]``, ``w = 1/|d|^2``; the node-diagonal block is
    ``[[d0,.1,.1],[.1,d1,.1],[.1,.1,d2]]`` with ``d_r = 1 + sum |other entries of row r|``
    (strictly diagonally dominant => SPD)."""
    nn = g ** 3
    n = 3 * nn
    v = np.arange(nn, dtype=np.int64).reshape(g, g, g)
    bmat = np.array([[1.0, 0.1, 0.1], [0.1, 1.0, 0.1], [0.1, 0.1, 1.0]])
    wsum = np.zeros((g, g, g))
    rows, cols, vals = [], [], []
    offsets = [o for o in itertools.product((-1, 0, 1), repeat=3) if o > (0, 0, 0)]
    for o in offsets:
        w = 1.0 / sum(d * d for d in o)
        hi, lo = [], []
        for d in o:
            if d == 1:
                hi.append(slice(1, None)); lo.append(slice(None, -1))
            elif d == -1:
                hi.append(slice(None, -1)); lo.append(slice(1, None))
            else:
                hi.append(slice(None)); lo.append(slice(None))
        hi, lo = tuple(hi), tuple(lo)
        big, small = v[hi].ravel(), v[lo].ravel()      # big = small + offset > small
        wsum[hi] += w
        wsum[lo] += w
        for a in range(3):
            for b in range(3):
                rows.append(3 * big + a)
                cols.append(3 * small + b)
                vals.append(np.full(big.size, -w * bmat[a, b]))
    u = v.ravel()
    delta = 1.2 + 1.2 * wsum.ravel()
    for a in range(3):
        rows.append(3 * u + a); cols.append(3 * u + a); vals.append(delta)
        for b in range(a):
            rows.append(3 * u + a); cols.append(3 * u + b); vals.append(np.full(nn, 0.1))
    return n, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), True


GENERATORS = {
    "lap2d": laplace2d,
    "lap3d": laplace3d,
    "cd3d": convdiff3d,
    "el3d": elasticity3d,
}

# BASELINE.json configs -> (generator, grid, factorization kind)
CONFIGS = {
    "cfg1": ("lap2d", 1000, "cholesky"),
    "cfg2": ("lap3d", 100, "cholesky"),
    "cfg3": ("el3d", 80, "cholesky"),
    "cfg4": ("cd3d", 100, "lu"),
    "cfg5": ("lap3d", 160, "cholesky"),
}


def generate(kind: str, g: int):
    return GENERATORS[kind](g)


def generate_complex(kind: str, g: int, hermitian: bool, amplitude: float = 0.45, diag_phase=None):
    """A complex matrix on the pattern of generate(kind, g): every entry (i, j) of the real matrix is rotated by the
    closed-form phase theta = amplitude * sin(1.3 i + 0.7 j + 0.5) (no RNG).  hermitian=True keeps the diagonal real, so that
    the lower triangle of a "symmetric" pattern is the lower triangle of a Hermitian positive definite matrix (diagonal
    dominance is a statement about moduli) -- the input of the reference's zpotrf path; hermitian=False also rotates the
    diagonal, by theta or, if given, by exactly diag_phase (pi/2: a purely IMAGINARY dominant diagonal) -- for the LU path,
    where a "symmetric" pattern then means A = A^T."""
    n, rows, cols, vals, symmetric = GENERATORS[kind](g)
    theta = amplitude * np.sin(1.3 * rows + 0.7 * cols + 0.5)
    diag = rows == cols
    if hermitian:
        theta[diag] = 0.0
    elif diag_phase is not None:
        theta[diag] = diag_phase
    return n, rows, cols, vals * np.exp(1j * theta), symmetric


def write_matrix_market(path, n, rows, cols, vals, symmetric):
    """MatrixMarket coordinate text, the only input format of the reference's driver.  Complex values are written with the
    field "complex" and two numbers per entry (the reader takes every field other than "real" as complex, CH.c:440-445)."""
    cplx = np.iscomplexobj(vals)
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate %s %s\n" % ("complex" if cplx else "real",
                                                               "symmetric" if symmetric else "general"))
        f.write("%d %d %d\n" % (n, n, len(rows)))
        chunk = 5_000_000
        for i in range(0, len(rows), chunk):
            v = vals[i:i + chunk]
            cols_ = [rows[i:i + chunk] + 1, cols[i:i + chunk] + 1] + ([v.real, v.imag] if cplx else [v])
            np.savetxt(f, np.column_stack(cols_), fmt="%d %d %.17g" + (" %.17g" if cplx else ""))


def write_binary(path, n, rows, cols, vals, symmetric):
    """Binary triplet container of the drop-in library (dropin/matrix_reader.hpp): magic "SFB200T1", int64 flags / nrow /
    ncol / nz, then Ti, Tj (0-based int64) and Tx (float64) in file order.  SparseFrame_read_matrix recognises it by
    the magic and builds the same Cp/Ci/Cx as from the MatrixMarket text of the same triplets (explicit zeros are
    dropped here, as the text reader drops them, CH.c:493)."""
    cplx = np.iscomplexobj(vals)             # flag bit 1; the values are then (re, im) pairs
    rows, cols = np.asarray(rows, np.int64), np.asarray(cols, np.int64)
    vals = np.asarray(vals, np.complex128 if cplx else np.float64)
    keep = vals != 0.0
    rows, cols, vals = rows[keep], cols[keep], vals[keep]
    with open(path, "wb") as f:
        f.write(b"SFB200T1")
        np.array([(1 if symmetric else 0) | (2 if cplx else 0), n, n, len(rows)], np.int64).tofile(f)
        rows.tofile(f)
        cols.tofile(f)
        vals.tofile(f)
