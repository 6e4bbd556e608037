"""Exact-rational basis enumerator (TEST INFRASTRUCTURE, pure Python, tiny LPs).

Independent of the floating-point oracle: Gauss-Jordan over fractions.Fraction,
so a basis is singular iff its determinant is exactly 0 and x_B, z are exact.
Used to pin oracle/enumcpu.c (and through it the CUDA kernels) on the goldens
of tests/golden/ and on random small integer LPs.

Semantics restated from the reference: gather Canonical.cpp:183-187; feasible
iff every x_i >= -1e-9 (Canonical.cpp:165-177); z = c.x (Canonical.cpp:79-87);
subsets in itertools.combinations (= lexicographic) order; strict improvement
keeps the lowest rank.
"""
from fractions import Fraction
from itertools import combinations

EPS = Fraction(1, 10**9)
FEASIBLE, INFEASIBLE, SINGULAR = 0, 1, 2


def solve_exact(cols, rhs):
    """Solve [cols] x = rhs exactly; cols is a list of m columns (lists of m). None if singular."""
    m = len(rhs)
    M = [[Fraction(cols[j][i]) for j in range(m)] + [Fraction(rhs[i])] for i in range(m)]
    for k in range(m):
        piv = next((r for r in range(k, m) if M[r][k] != 0), None)
        if piv is None:
            return None
        M[k], M[piv] = M[piv], M[k]
        pk = M[k][k]
        M[k] = [v / pk for v in M[k]]
        for r in range(m):
            if r != k and M[r][k] != 0:
                f = M[r][k]
                M[r] = [a - f * bb for a, bb in zip(M[r], M[k])]
    return [M[i][m] for i in range(m)]


def enumerate_exact(A, b, c, maximize):
    """A: m x n nested lists / array of exactly representable numbers.

    Returns dict(status=[per rank], z=[Fraction|None], x=[list|None], n_singular,
    n_infeasible, n_feasible, best_rank, best_basis, best_x, best_z,
    tied_ranks=[ranks attaining the optimum])."""
    m, n = len(A), len(A[0])
    Af = [[Fraction(A[i][j]) for j in range(n)] for i in range(m)]
    bf = [Fraction(v) for v in b]
    cf = [Fraction(v) for v in c]
    out = dict(status=[], z=[], x=[], n_singular=0, n_infeasible=0, n_feasible=0,
               best_rank=None, best_basis=None, best_x=None, best_z=None, tied_ranks=[])
    best_key = None
    for rank, S in enumerate(combinations(range(n), m)):
        x = solve_exact([[Af[i][j] for i in range(m)] for j in S], bf)
        if x is None:
            out["status"].append(SINGULAR); out["z"].append(None); out["x"].append(None)
            out["n_singular"] += 1
            continue
        z = sum(cf[j] * xi for j, xi in zip(S, x))
        out["z"].append(z); out["x"].append(x)
        if any(xi < -EPS for xi in x):
            out["status"].append(INFEASIBLE); out["n_infeasible"] += 1
            continue
        out["status"].append(FEASIBLE); out["n_feasible"] += 1
        key = -z if maximize else z
        if best_key is None or key < best_key:
            best_key = key
            out.update(best_rank=rank, best_basis=list(S), best_x=x, best_z=z, tied_ranks=[rank])
        elif key == best_key:
            out["tied_ranks"].append(rank)
    return out
