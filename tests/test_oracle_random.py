"""CPU: oracle vs the exact-rational enumerator on random small integer LPs
(hypothesis), vs SciPy HiGHS on the dense synthetic LPs, rank/unrank, ranges."""
from itertools import combinations

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import exact
from simplexmethod_b200 import lpgen


@st.composite
def small_lp(draw):
    m = draw(st.integers(1, 4))
    n = draw(st.integers(m, min(m + 5, 8)))
    ints = st.integers(-4, 4)
    A = [[draw(ints) for _ in range(n)] for _ in range(m)]
    b = [draw(st.integers(-6, 6)) for _ in range(m)]
    c = [draw(ints) for _ in range(n)]
    return A, b, c, draw(st.booleans())


@settings(max_examples=150, deadline=None)
@given(small_lp())
def test_oracle_vs_exact_random(oracle, lp):
    A, b, c, mx = lp
    m = len(A)
    if not any(any(row) for row in A):
        return                                       # all-zero A: scale 0, everything singular both ways
    ex = exact.enumerate_exact(A, b, c, mx)
    res, status = oracle.solve(np.array(A, dtype=float), b, c, mx, want_status=True)
    assert list(status) == ex["status"]
    assert (res.n_singular, res.n_infeasible, res.n_feasible) == (ex["n_singular"], ex["n_infeasible"], ex["n_feasible"])
    if ex["best_rank"] is None:
        assert res.status == 1
    else:
        # rounding may reorder exact ties, but only among the tied ranks
        assert res.best_rank in ex["tied_ranks"]
        assert res.objective == pytest.approx(float(ex["best_z"]), rel=1e-9, abs=1e-9)


@pytest.mark.parametrize("n,m", [(9, 4), (12, 1), (7, 7), (40, 12), (64, 16)])
def test_rank_unrank_roundtrip(oracle, n, m):
    import ctypes as C
    L = oracle.lib()
    total = L.enumcpu_binomial(n, m)
    assert total > 0
    S = (C.c_int32 * m)()
    probes = range(total) if total <= 200 else [0, 1, 2, total // 3, total // 2, total - 2, total - 1]
    combos = list(combinations(range(n), m)) if total <= 200 else None
    for r in probes:
        assert L.enumcpu_unrank(n, m, r, S) == 0
        assert L.enumcpu_rank(n, m, S) == r
        if combos:
            assert tuple(S) == combos[r]
    assert L.enumcpu_unrank(n, m, total, S) != 0


def test_range_merge_equals_full(oracle):
    A, b, c, mx = lpgen.dense_lp(5, 12, 7)
    full, _ = oracle.solve(A, b, c, mx)
    cuts = [0, 1, 100, 101, 500, 792]
    parts = [oracle.solve(A, b, c, mx, rank_begin=a, rank_end=e)[0] for a, e in zip(cuts[:-1], cuts[1:]) if a != e or a]
    assert sum(p.n_feasible for p in parts) == full.n_feasible
    assert sum(p.n_infeasible for p in parts) == full.n_infeasible
    best = min(((p.key, p.best_rank) for p in parts if p.status == 0))
    assert best == (full.key, full.best_rank)
    mt, _ = oracle.solve(A, b, c, mx, n_threads=5)
    assert (mt.best_rank, mt.n_feasible, mt.n_infeasible, mt.n_singular) == \
           (full.best_rank, full.n_feasible, full.n_infeasible, full.n_singular)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_dense_8_24_vs_highs(oracle, seed):
    """BASELINE config 2: independent optimum from SciPy HiGHS (same support, objective 1e-9)."""
    from scipy.optimize import linprog
    A, b, c, mx = lpgen.dense_lp(8, 24, seed)
    res, _ = oracle.solve(A, b, c, mx, n_threads=4)
    assert res.status == 0 and res.n_bases == 735471 and res.n_singular == 0
    h = linprog(c, A_eq=A, b_eq=b, bounds=(0, None), method="highs")
    assert h.status == 0
    assert res.objective == pytest.approx(h.fun, rel=1e-9)
    assert sorted(np.nonzero(h.x > 1e-9)[0].tolist()) == list(res.basis)[:8]
    # the generator's promise: basis {0..m-1} is feasible
    stt, x, _ = oracle.eval_basis(A, b, c, mx, list(range(8)))
    assert stt == oracle.FEASIBLE and min(x) >= 0.5 - 1e-9


def test_generator_is_pinned():
    """SplitMix64 stream and the LP built from it must never drift (goldens depend on it)."""
    g = lpgen.SplitMix64(1)
    assert [g.next() for _ in range(3)] == [0x910A2DEC89025CC1, 0xBEEB8DA1658EEC67, 0xF893A2EEFB32555E]
    A, b, c, mx = lpgen.dense_lp(3, 5, 42)
    assert A.flags.f_contiguous and not mx
    assert float(A[0, 0]).hex() == (2.0 * (lpgen.SplitMix64(42).next() >> 11) / 9007199254740992.0 - 1.0).hex()


def test_bad_arguments(oracle):
    from simplexmethod_b200 import _abi
    A, b, c, mx = lpgen.dense_lp(3, 6, 1)
    res, _ = oracle.solve(A, b, c, mx, rank_begin=5, rank_end=3)
    assert res.status == _abi.ERR_RANGE
    res, _ = oracle.solve(A, b, c, mx, rank_begin=0, rank_end=21)
    assert res.status == _abi.ERR_RANGE
    bad = A.copy(); bad[1, 1] = np.nan
    res, _ = oracle.solve(bad, b, c, mx)
    assert res.status == _abi.ERR_NONFINITE
