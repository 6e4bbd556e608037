"""BASELINE config 2: enumeration cross-checked against the reference's simplex
(restated in oracle/simplex_ref.py from src/SimplexSolover.h) — same optimal basis
and objective; plus the README's "compare SimplexSolver and EnumerationSolver"."""
import numpy as np
import pytest

from oracle import simplex_ref
from simplexmethod_b200 import lpgen


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_enumeration_vs_simplex_dense_8_24(oracle, seed):
    A, b, c, mx = lpgen.dense_lp(8, 24, seed)
    x, z, basis, iters = simplex_ref.solve_with_basis(A, b, c, list(range(8)), mx)
    res, _ = oracle.solve(A, b, c, mx, n_threads=4)
    assert basis == list(res.basis)[:8]
    assert z == pytest.approx(res.objective, rel=1e-9)
    assert iters < 200


def test_simplex_on_reference_fixtures(oracle):
    # input_symmetric.txt after ToCanonical: slack basis {3,4}; SURVEY A.1 optimum x=(5,0,0), z=35
    A, b, c, mx = lpgen.lab_symmetric_canonical()
    x, z, basis, _ = simplex_ref.solve_with_basis(A, b, c, [3, 4], mx, n_orig=3)
    assert x.tolist() == pytest.approx([5.0, 0.0, 0.0]) and z == pytest.approx(35.0) and basis == [0, 3]
    # src/main.cpp LP: optimum z=24
    A, b, c, mx = lpgen.main_cpp_canonical()
    x, z, basis, _ = simplex_ref.solve_with_basis(A, b, c, [3, 4], mx, n_orig=3)
    assert z == pytest.approx(24.0) and x.tolist() == pytest.approx([0.0, 0.0, 6.0])


@pytest.mark.gpu
@pytest.mark.parametrize("m,n,seed", [(8, 24, 1), (8, 24, 2), (8, 24, 3), (10, 30, 1), (12, 40, 1)])
def test_gpu_enumeration_vs_simplex(gpu_lib, m, n, seed):
    import simplexmethod_b200 as sm
    A, b, c, mx = lpgen.dense_lp(m, n, seed)
    x_s, z_s, basis_s, _ = simplex_ref.solve_with_basis(A, b, c, list(range(m)), mx)
    solver = sm.EnumerationSolver(sm.Canonical(A, b, c, list(range(m)), minimize=not mx))
    x_e = solver.solve()
    assert solver.optimalBasis() == basis_s
    assert solver.objective() == pytest.approx(z_s, rel=1e-9)
    assert np.allclose(x_e, x_s, rtol=1e-9, atol=1e-12)
