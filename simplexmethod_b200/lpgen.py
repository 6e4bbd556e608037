"""Deterministic synthetic LPs for the enumeration path.

All generators return canonical-form data ``(A, b, c, maximize)`` with ``A`` an
``(m, n)`` float64 array in COLUMN-MAJOR (Fortran) order — the memory layout of
``Eigen::MatrixXd`` that ``Canonical::GetConstraintsMatrix()`` hands out
(reference: src/ProblemTypes/Canonical.cpp:126-129).

The PRNG is SplitMix64; ``u = (next() >> 11) * 2**-53``.  Sums are evaluated
left to right with separate multiply and add (no FMA), in pure Python floats,
so the same bits come out on every machine.
"""
from __future__ import annotations

import numpy as np

_GAMMA = 0x9E3779B97F4A7C15
_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.state = seed & _M64

    def next(self) -> int:
        self.state = (self.state + _GAMMA) & _M64
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def uniform(self) -> float:
        return (self.next() >> 11) * (1.0 / 9007199254740992.0)


def dense_lp(m: int, n: int, seed: int):
    """Dense non-degenerate LP(m, n, seed): minimise, bounded, basis {0..m-1} feasible.

    Draw order: A column-major (2u-1); x0[i] = 0.5+u; y[i] = 2u-1; s[j] = 0.1+u.
    b = A[:, :m] @ x0 (so the first m columns are a feasible basis — the start a
    simplex cross-check needs); c = A' y + s (dual feasible => bounded below).
    """
    g = SplitMix64(seed)
    A = np.empty((m, n), dtype=np.float64, order="F")
    for j in range(n):
        for i in range(m):
            A[i, j] = 2.0 * g.uniform() - 1.0
    x0 = [0.5 + g.uniform() for _ in range(m)]
    y = [2.0 * g.uniform() - 1.0 for _ in range(m)]
    s = [0.1 + g.uniform() for _ in range(n)]
    b = np.empty(m)
    for i in range(m):
        acc = 0.0
        for j in range(m):
            acc = acc + float(A[i, j]) * x0[j]
        b[i] = acc
    c = np.empty(n)
    for j in range(n):
        acc = 0.0
        for i in range(m):
            acc = acc + float(A[i, j]) * y[i]
        c[j] = acc + s[j]
    return A, b, c, False


# Beale's cycling LP (1955), all data dyadic: min -3/4 x3 + 20 x4 - 1/2 x5 + 6 x6
_BEALE_A = [[1, 0, 0, 0.25, -8.0, -1.0, 9.0],
            [0, 1, 0, 0.50, -12.0, -0.5, 3.0],
            [0, 0, 1, 0.00, 0.0, 1.0, 0.0]]
_BEALE_B = [0.0, 0.0, 1.0]
_BEALE_C = [0.0, 0.0, 0.0, -0.75, 20.0, -0.5, 6.0]


def beale_lp():
    """Beale's 3x7 cycling LP in canonical form."""
    A = np.asfortranarray(np.array(_BEALE_A, dtype=np.float64))
    return A, np.array(_BEALE_B), np.array(_BEALE_C), False


def degenerate_lp():
    """Degenerate m=10, n=30 LP (BASELINE config 5): exact ties and singular bases.

    Columns 0..20: three Beale blocks on the diagonal (rows 0-2, 3-5, 6-8).
    Row 9 couples the blocks: x3 + x10 + x17 + x21 = 3 (column 21 is its slack).
    Columns 22..29 duplicate or rescale earlier columns — including columns of
    the optimal basis — with costs scaled alike, so many bases are exactly
    singular and the optimum is attained by several bases with equal objective;
    the lowest lexicographic rank must win.  All entries are dyadic rationals.
    """
    m, n = 10, 30
    A = np.zeros((m, n), dtype=np.float64, order="F")
    c = np.zeros(n)
    b = np.zeros(m)
    for blk in range(3):
        for i in range(3):
            for j in range(7):
                A[3 * blk + i, 7 * blk + j] = _BEALE_A[i][j]
            b[3 * blk + i] = _BEALE_B[i]
        for j in range(7):
            c[7 * blk + j] = _BEALE_C[j]
    for j in (3, 10, 17):
        A[9, j] = 1.0
    A[9, 21] = 1.0
    b[9] = 3.0
    # (source column, scale): duplicates and rescalings
    extra = [(0, 1.0), (3, 2.0), (5, 1.0), (7, 0.5), (12, 1.0), (14, 4.0), (21, 1.0), (10, 0.25)]
    for k, (src, sc) in enumerate(extra):
        A[:, 22 + k] = sc * A[:, src]
        c[22 + k] = sc * c[src]
    return A, b, c, False


def small_degenerate_lp():
    """Degenerate m=7, n=18 LP (31 824 bases) for fast sharding / sub-range tests: two Beale blocks (rows 0-2 and
    3-5), a coupling row x_B3 + x_C3 + slack = 2, and three duplicated / rescaled columns, two of them placed FIRST
    in the column order: columns 0 and 1 are identical, so every child task whose prefix starts (0, 1, .) is
    singular as a whole (the shared kernel books such subtrees in bulk), and column 2 = 2 x column 5.
    All entries dyadic; many exact ties at the optimum."""
    m = 7
    cols, cost = [], []

    def beale(blk, j, sc=1.0):
        v = np.zeros(m)
        for i in range(3):
            v[3 * blk + i] = sc * _BEALE_A[i][j]
        if j == 3:
            v[6] = sc * 1.0                      # coupling row
        return v, sc * _BEALE_C[j]

    order = [(0, 0, 1.0), (0, 0, 1.0), (0, 3, 2.0)] + [(0, j, 1.0) for j in range(1, 7)] + [(1, j, 1.0) for j in range(7)]
    for blk, j, sc in order:
        v, cj = beale(blk, j, sc)
        cols.append(v); cost.append(cj)
    slack = np.zeros(m); slack[6] = 1.0
    cols.append(slack); cost.append(0.0)
    v, cj = beale(1, 3, 0.5)
    cols.append(v); cost.append(cj)
    A = np.asfortranarray(np.array(cols).T)
    b = np.array([_BEALE_B[0], _BEALE_B[1], _BEALE_B[2], _BEALE_B[0], _BEALE_B[1], _BEALE_B[2], 2.0])
    assert A.shape == (7, 18)
    return A, b, np.array(cost), False


def lab_symmetric_canonical():
    """Config 1: the reference's input_symmetric.txt LP after Symmetrical::ToCanonical.

    max 7x1+8x2+3x3, x1+2x2+3x3<=10, 4x1+5x2+6x3<=20 (reference:
    input_symmetric.txt:1-8) -> [A | I], slack basis {3,4}, maximise
    (reference: src/ProblemTypes/Symmetrical.cpp:163-189).  n_orig = 3.
    """
    A = np.asfortranarray(np.array([[1, 2, 3, 1, 0], [4, 5, 6, 0, 1]], dtype=np.float64))
    return A, np.array([10.0, 20.0]), np.array([7.0, 8.0, 3.0, 0.0, 0.0]), True


def main_cpp_canonical():
    """The LP hard-coded in the reference demo (src/main.cpp:48-57) in canonical form."""
    A = np.asfortranarray(np.array([[1, 1, 1, 1, 0], [2, 1, 0, 0, 1]], dtype=np.float64))
    return A, np.array([6.0, 8.0]), np.array([3.0, 2.0, 4.0, 0.0, 0.0]), True


def test_canonical_fixture():
    """Fixture of the reference's tests/test_canonical.cpp:12-22 (minimise)."""
    A = np.asfortranarray(np.array([[1, 2, 1, 0], [3, 4, 0, 1]], dtype=np.float64))
    return A, np.array([5.0, 6.0]), np.array([7.0, 8.0, 0.0, 0.0]), False


def readme_shaped_lp():
    """A lab-assignment-shaped LP (reference README.md:5-8: 5 variables, 3
    inequalities + 1 equality) written for this repo, already canonical:
    m = 4 rows, n = 5 + 3 slacks = 8 columns, maximise.  Small integers, so the
    exact-rational enumerator can check every one of its C(8,4)=70 bases."""
    A = np.zeros((4, 8), order="F")
    A[0, :5] = [2, 1, 1, 0, 3]
    A[1, :5] = [1, 3, 0, 2, 1]
    A[2, :5] = [0, 1, 4, 1, 2]
    A[3, :5] = [1, 1, 1, 1, 1]          # equality row: no slack
    A[0, 5] = A[1, 6] = A[2, 7] = 1.0
    b = np.array([12.0, 15.0, 16.0, 7.0])
    c = np.array([3.0, 5.0, 4.0, 2.0, 6.0, 0.0, 0.0, 0.0])
    return A, b, c, True
