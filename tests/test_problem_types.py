"""the reference's problem types and parser around the enumeration path.

CPU: the reference's gtest cases restated on the Python mirror (tests/test_common.cpp,
test_symmetrical.cpp, test_canonical.cpp, test_transformations.cpp, test_parser.cpp), and
the conversions checked for MEANING: a general-form LP pushed through
Common -> ToSymmetrical -> ToCanonical and enumerated by the oracle has the optimum HiGHS
finds on the general form, and so has its Common::GetDual() (strong duality).
GPU: the C++ types (simplexmethod_b200/cpp) run the same flows through libenumgpu and
must reproduce the oracle on the Python-built canonical form bit for bit.
"""
import os
import subprocess

import numpy as np
import pytest

import simplexmethod_b200 as sm
from simplexmethod_b200 import Common, ConstraintType as CT, Symmetrical, SymmetricalParser, VariableType as VT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "simplexmethod_b200", "cpp")
A22, B2, C2 = np.array([[1.0, 2.0], [3.0, 4.0]]), np.array([5.0, 6.0]), np.array([7.0, 8.0])


def lab_common():
    """The general-form lab LP hard-coded in cpp/tests/enum_demo.cpp (--lab)."""
    A = np.array([[2, 1, 1, 1, -3], [1, 3, 2, 2, -1], [1, 1, 4, 1, -2], [1, 1, 1, 1, -1]], dtype=float)
    return Common(A, [12, 15, 16, 7], [3, 5, 1, 2, -6], [CT.LessOrEqual, CT.LessOrEqual, CT.GreaterOrEqual, CT.Equal],
                  [VT.NonNegative, VT.NonNegative, VT.NonNegative, VT.Free, VT.NonPositive], True)


def highs_general(com: Common):
    """Optimum of a general-form LP straight from its definition (no conversion)."""
    from scipy.optimize import linprog
    A, b, c = com.GetConstraintsMatrix(), com.GetRightHandSide(), com.GetObjectiveCoefficients()
    rows = com.GetConstraintTypes()
    le = [i for i, t in enumerate(rows) if t is CT.LessOrEqual]
    ge = [i for i, t in enumerate(rows) if t is CT.GreaterOrEqual]
    eq = [i for i, t in enumerate(rows) if t is CT.Equal]
    A_ub = np.vstack([A[le], -A[ge]]) if le or ge else None
    b_ub = np.concatenate([b[le], -b[ge]]) if le or ge else None
    bounds = [{VT.Free: (None, None), VT.NonNegative: (0, None), VT.NonPositive: (None, 0)}[t] for t in com.GetVariableTypes()]
    sign = -1.0 if com.IsMaximization() else 1.0
    r = linprog(sign * c, A_ub=A_ub, b_ub=b_ub, A_eq=A[eq] if eq else None, b_eq=b[eq] if eq else None,
                bounds=bounds, method="highs")
    return r.status, (sign * r.fun if r.status == 0 else None), r.x


def oracle_solve(oracle, can):
    res, _ = oracle.solve(can.GetConstraintsMatrix(), can.GetRightHandSide(), can.GetObjectiveCoefficients(),
                          can.IsMaximization())
    return res


# ---- the reference's gtest cases, restated ---------------------------------------------------

def test_common_cases():
    """tests/test_common.cpp:39-94."""
    com = Common(A22, B2, C2, [CT.LessOrEqual, CT.GreaterOrEqual], [VT.NonNegative, VT.NonNegative], True)
    assert com.IsMaximization() and com.GetConstraintsMatrix().shape == (2, 2)
    assert com.GetRightHandSide().size == 2 and com.GetObjectiveCoefficients().size == 2
    assert com.Evaluate([1, 2]) == 23.0
    with pytest.raises(ValueError):
        Common(A22, [1, 2, 3], C2, [CT.LessOrEqual, CT.GreaterOrEqual], [VT.NonNegative, VT.NonNegative], True)
    with pytest.raises(ValueError):
        Common(A22, B2, C2, [CT.LessOrEqual], [VT.NonNegative, VT.NonNegative], True)
    dual = com.GetDual()
    assert not dual.IsMaximization() and dual.GetConstraintsMatrix().shape == (2, 2)
    assert dual.GetVariableTypes() == [VT.NonNegative, VT.NonPositive]
    assert dual.GetConstraintTypes() == [CT.GreaterOrEqual, CT.GreaterOrEqual]
    back = dual.GetDual()
    assert back.IsMaximization() and (back.GetConstraintsMatrix() == A22).all()
    assert back.GetConstraintTypes() == com.GetConstraintTypes() and back.GetVariableTypes() == com.GetVariableTypes()


def test_symmetrical_cases():
    """tests/test_symmetrical.cpp:27-97."""
    sym = Symmetrical(A22, B2, C2, True)
    assert sym.IsMaximization() and sym.GetConstraintsMatrix().shape == (2, 2)
    dual = sym.GetDual()
    assert not dual.IsMaximization() and dual.GetConstraintsMatrix().shape == (2, 2)
    assert (dual.GetConstraintsMatrix() == A22.T).all()
    assert (dual.GetRightHandSide() == C2).all() and (dual.GetObjectiveCoefficients() == B2).all()
    can = sym.ToCanonical()
    assert can.GetConstraintsMatrix().shape == (2, 4) and can.GetBasisIndices() == [2, 3] and can.IsMaximization()
    assert can.GetOriginalVariablesCount() == 2
    cmin = Symmetrical(A22, B2, C2, False).ToCanonical()
    assert cmin.GetConstraintsMatrix().shape == (2, 6) and cmin.GetBasisIndices() == [4, 5] and not cmin.IsMaximization()
    assert (cmin.GetConstraintsMatrix()[:, 2:4] == -np.eye(2)).all() and (cmin.GetObjectiveCoefficients()[2:] == 0).all()
    com = sym.ToCommon()
    assert com.IsMaximization() and len(com.GetConstraintTypes()) == 2 and len(com.GetVariableTypes()) == 2
    assert com.GetConstraintTypes() == [CT.LessOrEqual] * 2


def test_canonical_to_other_forms():
    """tests/test_canonical.cpp:78-89 and Canonical.cpp:230-303."""
    A, b, c, _ = sm.lpgen.test_canonical_fixture()
    can = sm.Canonical(A, b, c, [2, 3], True)
    can.SetOriginalVariablesCount(2)
    com = can.ToCommon()
    assert com.GetObjectiveCoefficients().size == 2 and com.GetConstraintsMatrix().shape == (2, 2)
    assert com.GetConstraintTypes() == [CT.Equal] * 2 and not com.IsMaximization()
    sym = can.ToSymmetrical()
    assert not sym.IsMaximization()
    assert sym.GetConstraintsMatrix().tolist() == [[1, 2], [-1, -2], [3, 4], [-3, -4]]
    assert sym.GetRightHandSide().tolist() == [5, -5, 6, -6]


def test_transformations_cases():
    """tests/test_transformations.cpp:6-61."""
    com = Common(A22, B2, C2, [CT.LessOrEqual] * 2, [VT.NonNegative] * 2, True)
    sym = com.ToSymmetrical()
    assert (sym.GetConstraintsMatrix() == A22).all()
    assert sym.ToCanonical().GetConstraintsMatrix().shape[0] == 2
    dd = Symmetrical(A22, B2, C2, True).GetDual().GetDual()
    assert dd.IsMaximization() and (dd.GetConstraintsMatrix() == A22).all()
    assert (dd.GetRightHandSide() == B2).all() and (dd.GetObjectiveCoefficients() == C2).all()


def test_every_row_and_variable_kind():
    """Common.cpp:169-348 on one LP with all six kinds; same expected matrix as cpp/tests/host_tests.cpp."""
    gen = Common(np.arange(1.0, 10.0).reshape(3, 3), [10, 11, 12], [1, -2, 3],
                 [CT.LessOrEqual, CT.GreaterOrEqual, CT.Equal], [VT.Free, VT.NonNegative, VT.NonPositive], False)
    sym = gen.ToSymmetrical()
    assert sym.IsMaximization()
    assert sym.GetConstraintsMatrix().tolist() == [[1, -1, 2, -3], [-4, 4, -5, 6], [7, -7, 8, -9], [-7, 7, -8, 9]]
    assert sym.GetRightHandSide().tolist() == [10, -11, 12, -12]
    assert sym.GetObjectiveCoefficients().tolist() == [-1, 1, 2, 3]
    can = gen.ToCanonical()
    assert can.GetConstraintsMatrix().shape == (4, 8) and can.GetOriginalVariablesCount() == 4
    dual = gen.GetDual()
    assert dual.IsMaximization() and dual.GetVariableTypes() == [VT.NonPositive, VT.NonNegative, VT.Free]
    assert dual.GetConstraintTypes() == [CT.Equal, CT.LessOrEqual, CT.GreaterOrEqual]


def test_parser_cases(tmp_path):
    """tests/test_parser.cpp:4-81 and the reference's input_symmetric.txt."""
    p = SymmetricalParser()
    s1 = p.ParseFromString("\n  maximize\n\n objective:\n 3 5\n\n constraints:\n 1 2 10\n 3 4 20\n")
    assert s1 is not None and s1.IsMaximization() and s1.GetConstraintsMatrix().shape == (2, 2)
    assert s1.GetRightHandSide().tolist() == [10, 20]
    s2 = p.ParseFromString("minimize\nobjective:\n7 8\nsubject to:\n1 1 5\n2 3 12\n")
    assert s2 is not None and not s2.IsMaximization()
    s3 = p.ParseFromString("# c\nmaximize\nobjective:\n1 2 3  # more\nconstraints:\n1 0 0 5 # a\r\n0 1 0 6\n0 0 1 7\n")
    assert s3 is not None and s3.GetObjectiveCoefficients().size == 3
    assert p.ParseFromString("maximize\n# nothing else\n") is None and p.GetLastError()
    assert p.ParseFromString("1 2 3\n") is None
    assert p.ParseFromString("max\nobjective:\n1 2\nconstraints:\n1 2 3 4\n") is None
    assert p.ParseFromFile(str(tmp_path / "missing.txt")) is None
    lab = p.ParseFromFile(os.path.join(ROOT, "tests", "golden", "lab_lp_symmetric.txt"))
    A, b, c, mx = sm.lpgen.lab_symmetric_canonical()
    can = lab.ToCanonical()
    assert (can.GetConstraintsMatrix() == A).all() and (can.GetRightHandSide() == b).all()
    assert (can.GetObjectiveCoefficients() == c).all() and can.IsMaximization() == mx


# ---- the conversions keep the LP's meaning (oracle enumeration vs HiGHS on the general form) --

def test_lab_lp_primal_dual_vs_highs(oracle):
    com = lab_common()
    st, z, x = highs_general(com)
    assert st == 0
    rp = oracle_solve(oracle, com.ToCanonical())
    rd = oracle_solve(oracle, com.GetDual().ToCanonical())
    assert rp.status == 0 and rd.status == 0
    assert rp.objective == pytest.approx(z, rel=1e-9)
    # both are "max" problems in symmetric form: the min-dual has its costs negated (Common.cpp:342-345)
    assert -rd.objective == pytest.approx(z, rel=1e-9)


@pytest.mark.parametrize("seed", range(12))
def test_random_general_lps_vs_highs(oracle, seed):
    rng = np.random.default_rng(seed)
    m, n = int(rng.integers(2, 5)), int(rng.integers(2, 5))
    A = rng.integers(-4, 5, size=(m, n)).astype(float)
    x0 = rng.integers(-3, 4, size=n).astype(float)
    vts = [VT(int(k)) for k in rng.integers(0, 3, size=n)]
    x0 = np.array([abs(v) if t is VT.NonNegative else -abs(v) if t is VT.NonPositive else v for v, t in zip(x0, vts)])
    cts = [CT(int(k)) for k in rng.integers(0, 3, size=m)]
    slack = rng.integers(0, 3, size=m).astype(float)
    b = A @ x0 + np.array([s if t is CT.LessOrEqual else -s if t is CT.GreaterOrEqual else 0.0 for s, t in zip(slack, cts)])
    c = rng.integers(-5, 6, size=n).astype(float)
    com = Common(A, b, c, cts, vts, bool(rng.integers(0, 2)))
    st, z, _ = highs_general(com)
    can = com.ToCanonical()
    res = oracle_solve(oracle, can)
    if st == 0:
        # symmetric form is always "max": a min problem has its costs negated
        got = res.objective if com.IsMaximization() else -res.objective
        assert res.status == 0 and got == pytest.approx(z, rel=1e-9, abs=1e-9)
        rd = oracle_solve(oracle, com.GetDual().ToCanonical())
        gotd = rd.objective if com.GetDual().IsMaximization() else -rd.objective
        assert rd.status == 0 and gotd == pytest.approx(z, rel=1e-9, abs=1e-9)      # strong duality
    elif st == 2:
        assert res.status == 1                                   # infeasible: no feasible basis
    # st == 3 (unbounded): enumeration of vertices cannot detect it; nothing to compare


# ---- GPU: the C++ types through libenumgpu == oracle on the Python-built canonical form ------

@pytest.fixture(scope="module")
def built():
    subprocess.check_call(["make", "-C", CPP, "-s"])
    return os.path.join(CPP, "build")


def _facts(built, flag):
    out = subprocess.run([os.path.join(built, "enum_demo"), flag], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return {l.split()[0]: l.split()[1:] for l in out.stdout.splitlines()}


def _same_as_oracle(facts, tag, oracle, can):
    res = oracle_solve(oracle, can)
    m = can.GetConstraintsMatrix().shape[0]
    x = np.zeros(can.GetObjectiveCoefficients().size)
    for j, v in zip(list(res.basis)[:m], list(res.x_B)[:m]):
        x[j] = v
    assert [float(v) for v in facts[tag + "_x"]] == x[:can.GetOriginalVariablesCount()].tolist()
    assert float(facts[tag + "_objective"][0]) == res.objective
    assert [int(v) for v in facts[tag + "_basis"]] == list(res.basis)[:m]
    assert [int(v) for v in facts[tag + "_counts"]] == [res.n_bases, res.n_singular, res.n_infeasible, res.n_feasible]
    return res


@pytest.mark.gpu
def test_main_cpp_flow_cpp(built, gpu_lib, oracle):
    """src/main.cpp:48-113 with EnumerationSolver in place of Solver: z = 24 at x = (0,0,6)."""
    facts = _facts(built, "--main")
    com = Common([[1, 1, 1], [2, 1, 0]], [6, 8], [3, 2, 4], [CT.LessOrEqual] * 2, [VT.NonNegative] * 3, True)
    sym = com.ToSymmetrical()
    assert facts["dual_shape"] == ["3", "2", "0"]
    assert [float(v) for v in facts["initial_x"]] == [0, 0, 0, 6, 8] and facts["initial_basis_feasible"] == ["1"]
    assert float(facts["initial_z"][0]) == 0.0
    res = _same_as_oracle(facts, "primal", oracle, sym.ToCanonical())
    assert res.objective == 24.0 and [float(v) for v in facts["primal_x"]] == [0.0, 0.0, 6.0]
    rd = _same_as_oracle(facts, "dual", oracle, com.GetDual().ToCanonical())
    assert -rd.objective == pytest.approx(24.0, rel=1e-12)            # strong duality (min dual in max/<= form)
    # the min-form canonical of the symmetric dual carries zero-cost artificial columns (Symmetrical.cpp:191-222):
    # enumerating it is a relaxation, optimum 0 with the artificials basic
    assert oracle_solve(oracle, sym.GetDual().ToCanonical()).objective == 0.0


@pytest.mark.gpu
def test_lab_common_flow_cpp(built, gpu_lib, oracle):
    """General-form lab LP (README.md:5-8 shape): primal and Common::GetDual() through the C++ types on the GPU."""
    facts = _facts(built, "--lab")
    com = lab_common()
    rp = _same_as_oracle(facts, "primal", oracle, com.ToCanonical())
    rd = _same_as_oracle(facts, "dual", oracle, com.GetDual().ToCanonical())
    st, z, x = highs_general(com)
    assert st == 0 and rp.objective == pytest.approx(z, rel=1e-9) and -rd.objective == pytest.approx(z, rel=1e-9)
    xs = [float(v) for v in facts["primal_x"]]           # (x1, x2, x3, x4', x4'', x5') of the symmetric form
    assert com.Evaluate([xs[0], xs[1], xs[2], xs[3] - xs[4], -xs[5]]) == pytest.approx(z, rel=1e-9)


@pytest.mark.gpu
def test_python_flow_parse_to_solve(gpu_lib, oracle):
    """ParseFromFile -> ToCanonical -> EnumerationSolver.solve() in Python (config 1): x = (5,0,0), z = 35."""
    sym = SymmetricalParser().ParseFromFile(os.path.join(ROOT, "tests", "golden", "lab_lp_symmetric.txt"))
    solver = sm.EnumerationSolver(sym.ToCanonical())
    assert solver.solve().tolist() == [5.0, 0.0, 0.0] and solver.objective() == 35.0
    com = lab_common()
    for can in (com.ToCanonical(), com.GetDual().ToCanonical()):
        s = sm.EnumerationSolver(can)
        s.solve()
        res = oracle_solve(oracle, can)
        assert s.objective() == res.objective and s.bestRank() == res.best_rank
        assert [s.singularCount(), s.infeasibleCount(), s.feasibleCount()] == [res.n_singular, res.n_infeasible, res.n_feasible]
