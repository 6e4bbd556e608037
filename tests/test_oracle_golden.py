"""CPU: the oracle (oracle/enumcpu.c) against the exact-rational goldens of
tests/golden/tiny_lps.json and against the figures quoted in SURVEY.md App. A —
the reference's own fixtures: input_symmetric.txt, src/main.cpp:48-57,
tests/test_canonical.cpp:12-22 (incl. the EXPECT_DOUBLE_EQ pin :52-57)."""
import json
import os
from fractions import Fraction

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "tiny_lps.json")) as f:
    GOLD = json.load(f)

# independent of the JSON: numbers stated in SURVEY.md Appendix A
SURVEY_A = {
    "lab_symmetric": dict(counts=(0, 3, 7), rank=2, basis=[0, 3], z=35.0, x_head=[5.0, 0.0, 0.0]),
    "main_cpp": dict(counts=(1, 3, 6), rank=8, basis=[2, 4], z=24.0, x_head=[0.0, 0.0, 6.0]),
    "test_canonical": dict(counts=(0, 3, 3), rank=5, basis=[2, 3], z=0.0, x_head=[0.0, 0.0, 5.0, 6.0]),
    "beale": dict(counts=(10, 6, 19), rank=10, basis=[0, 3, 5], z=-1.25, x_head=None),
}


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_matches_exact_golden(oracle, name):
    g = GOLD[name]
    A = np.array(g["A"]); m, n = A.shape
    res, status = oracle.solve(A, g["b"], g["c"], g["maximize"], want_status=True)
    assert list(status) == g["status"]                      # per-basis classification
    assert (res.n_singular, res.n_infeasible, res.n_feasible) == (g["n_singular"], g["n_infeasible"], g["n_feasible"])
    assert res.n_bases == len(g["status"])
    assert res.best_rank == g["best_rank"]
    assert list(res.basis)[:m] == g["best_basis"]
    assert res.objective == pytest.approx(float(Fraction(g["best_z"])), rel=1e-12, abs=1e-12)
    for got, want in zip(list(res.x_B)[:m], g["best_x"]):
        assert got == pytest.approx(float(Fraction(want)), rel=1e-12, abs=1e-12)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_per_basis_values(oracle, name):
    """x_B and z of every non-singular basis within 1e-9 relative of the exact values."""
    from itertools import combinations
    g = GOLD[name]
    A = np.array(g["A"]); m, n = A.shape
    for rank, S in enumerate(combinations(range(n), m)):
        st, x, z = oracle.eval_basis(A, g["b"], g["c"], g["maximize"], S)
        assert st == g["status"][rank]
        if st == oracle.SINGULAR:
            continue
        assert z == pytest.approx(float(Fraction(g["z"][rank])), rel=1e-9, abs=1e-9)
        for got, want in zip(x, g["x"][rank]):
            assert got == pytest.approx(float(Fraction(want)), rel=1e-9, abs=1e-9)


@pytest.mark.parametrize("name", sorted(SURVEY_A))
def test_survey_appendix_a(oracle, name):
    g, s = GOLD[name], SURVEY_A[name]
    A = np.array(g["A"]); m, n = A.shape
    res, _ = oracle.solve(A, g["b"], g["c"], g["maximize"])
    assert (res.n_singular, res.n_infeasible, res.n_feasible) == s["counts"]
    assert res.best_rank == s["rank"] and list(res.basis)[:m] == s["basis"]
    assert res.objective == s["z"]
    if s["x_head"] is not None:
        x = np.zeros(n)
        for i in range(m):
            x[res.basis[i]] = res.x_B[i]
        assert x[: g["n_orig"]].tolist() == s["x_head"]


def test_reference_identity_basis_pin(oracle):
    """tests/test_canonical.cpp:41-58: basis {2,3} (identity) => x = (0,0,5,6), EXPECT_DOUBLE_EQ; :60-66 feasible."""
    g = GOLD["test_canonical"]
    st, x, z = oracle.eval_basis(np.array(g["A"]), g["b"], g["c"], False, [2, 3])
    assert st == oracle.FEASIBLE and x == [5.0, 6.0] and z == 0.0


def test_beale_singular_and_tied_ranks(oracle):
    g = GOLD["beale"]
    st = g["status"]
    assert [r for r, s in enumerate(st) if s == 2] == [1, 2, 4, 9, 11, 13, 19, 21, 23, 32]
    assert [r for r, s in enumerate(st) if s == 1] == [12, 14, 20, 22, 31, 34]
    zero = [r for r, s in enumerate(st) if s == 0 and Fraction(g["z"][r]) == 0]
    assert zero == [0, 5, 6, 7, 8, 15, 16, 17, 18, 25, 26, 27, 28, 29, 30]
