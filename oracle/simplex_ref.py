"""Revised simplex restated from the reference (TEST INFRASTRUCTURE, numpy).

The cross-check BASELINE.json's config 2 asks for ("cross-checked vs
SimplexSolver") and the README's "compare SimplexSolver and EnumerationSolver"
deliverable (reference README.md:40-42).  Follows the reachable part of the
reference's Solver:
  computeBFS        src/SimplexSolover.h:117-133   (B^-1 explicitly, x_B = B^-1 b,
                                                    singular basis -> error)
  simplexIter       src/SimplexSolover.h:135-209   (largest-coefficient pricing with
                                                    EPS guards, ratio test with EPS,
                                                    eta update of B^-1)
  solveWithBasis    src/SimplexSolover.h:408-451   (refactorise every iteration,
                                                    MAX_ITER = 10000, x.head(n_orig))
Two-phase / redundant-row recovery are unreachable in the reference
and are not restated: the caller supplies a feasible starting basis.
"""
import numpy as np

EPS = 1e-9          # SimplexSolover.h:13
MAX_ITER = 10000    # SimplexSolover.h:426


def _compute_bfs(A, b, N):
    B = A[:, N]
    if np.linalg.matrix_rank(B) < B.shape[0]:
        raise RuntimeError("Singular basis matrix")          # :125-126
    Binv = np.linalg.inv(B)
    x = np.zeros(A.shape[1])
    x[N] = Binv @ b
    return x, Binv


def _simplex_iter(A, b, c, N, Binv, maximize):
    m, n = A.shape
    yT = c[N] @ Binv
    L = [j for j in range(n) if j not in set(N.tolist())]     # complement(), :95-108
    enter = -1
    if maximize:
        best = -np.inf
        for j in L:
            d = c[j] - yT @ A[:, j]
            if d > best + EPS:
                best, enter = d, j
        if best <= EPS:
            return "optimal"
    else:
        best = np.inf
        for j in L:
            d = c[j] - yT @ A[:, j]
            if d < best - EPS:
                best, enter = d, j
        if best >= -EPS:
            return "optimal"
    u = Binv @ A[:, enter]
    xB = Binv @ b
    if np.all(u <= EPS):
        return "unbounded"
    theta, leave = np.inf, -1
    for i in range(m):
        if u[i] > EPS:
            r = xB[i] / u[i]
            if r < theta - EPS:
                theta, leave = r, i
    if leave == -1:
        return "unbounded"
    N[leave] = enter
    F = np.eye(m)
    for i in range(m):
        if i != leave:
            F[i, leave] = -u[i] / u[leave]
    F[leave, leave] = 1.0 / u[leave]
    Binv[:] = F @ Binv
    return "iter"


def solve_with_basis(A, b, c, basis, maximize, n_orig=None):
    """Returns (x[:n_orig], objective, final basis sorted, iterations)."""
    A = np.asarray(A, dtype=float); b = np.asarray(b, dtype=float); c = np.asarray(c, dtype=float)
    N = np.array(basis, dtype=int)
    x, Binv = _compute_bfs(A, b, N)
    for it in range(MAX_ITER):
        status = _simplex_iter(A, b, c, N, Binv, maximize)
        if status == "optimal":
            x, Binv = _compute_bfs(A, b, N)
            k = A.shape[1] if n_orig is None else n_orig
            return x[:k].copy(), float(c @ x), sorted(N.tolist()), it
        if status == "unbounded":
            raise RuntimeError("objective is unbounded")       # :443
        x, Binv = _compute_bfs(A, b, N)                         # :446
    raise RuntimeError("iteration limit reached")               # :450
