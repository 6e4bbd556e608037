"""C++ host side (simplexmethod_b200/cpp): the reference's problem types, parser
and the EnumerationSolver adapter over the C ABI."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "simplexmethod_b200", "cpp")
LAB = os.path.join(ROOT, "tests", "golden", "lab_lp_symmetric.txt")


@pytest.fixture(scope="module")
def built():
    subprocess.check_call(["make", "-C", CPP, "-s"])
    return os.path.join(CPP, "build")


def test_host_types_and_parser_cpu(built):
    """Restated reference gtest cases for Canonical / Symmetrical / parser; no GPU needed."""
    out = subprocess.run([os.path.join(built, "host_tests")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "all passed" in out.stdout


def test_adapter_links_only_the_c_abi(built):
    """The demo's undefined enumgpu symbols are all declared in include/enumgpu.h."""
    syms = subprocess.check_output(["nm", "-u", os.path.join(built, "enum_demo")], text=True)
    used = sorted({l.split()[-1].split("@")[0] for l in syms.splitlines() if "enumgpu_" in l})
    assert used == ["enumgpu_create", "enumgpu_destroy", "enumgpu_eval_basis", "enumgpu_last_error", "enumgpu_solve_hv"]
    header = open(os.path.join(ROOT, "include", "enumgpu.h")).read()
    assert all(name + "(" in header for name in used)


@pytest.mark.gpu
def test_reference_flow_config1_cpp(built, gpu_lib):
    """ParseFromFile(input_symmetric LP) -> ToCanonical -> EnumerationSolver.solve(): x=(5,0,0), z=35."""
    out = subprocess.run([os.path.join(built, "enum_demo"), LAB], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    facts = {l.split()[0]: l.split()[1:] for l in out.stdout.splitlines()}
    assert facts["initial_basis_feasible"] == ["1"]
    assert [float(v) for v in facts["initial_x"]] == [0, 0, 0, 10, 20]      # slack basis {3,4}
    assert [float(v) for v in facts["x"]] == [5.0, 0.0, 0.0]
    assert float(facts["objective"][0]) == 35.0
    assert facts["basis"] == ["0", "3"] and facts["best_rank"] == ["2"]
    assert facts["counts"] == ["10", "0", "3", "7"]


@pytest.mark.gpu
def test_python_canonical_per_basis_methods(gpu_lib, oracle):
    """Canonical.GetBasicSolution / IsFeasibleBasis on the GPU: the reference's identity-basis pin
    (tests/test_canonical.cpp:41-66) and unsorted / infeasible / singular bases vs the oracle."""
    import simplexmethod_b200 as sm
    from simplexmethod_b200 import lpgen
    A, b, c, _ = lpgen.test_canonical_fixture()
    can = sm.Canonical(A, b, c, [2, 3])
    assert can.GetBasicSolution().tolist() == [0.0, 0.0, 5.0, 6.0] and can.IsFeasibleBasis()
    assert not sm.Canonical(A, b, c, [0, 3]).IsFeasibleBasis()               # infeasible
    for basis, want in (([2, 1], 0), ([3, 1], 1)):                           # unsorted order; feasible / infeasible
        can2 = sm.Canonical(A, b, c, basis)
        x = can2.GetBasicSolution()                                          # returned whether feasible or not
        st, xo, _ = oracle.eval_basis(A[:, basis], b, c[basis], False, [0, 1])
        assert st == want and [x[basis[0]], x[basis[1]]] == xo
        assert can2.IsFeasibleBasis() == (want == 0)
    A2, b2, c2, mx = lpgen.main_cpp_canonical()
    with pytest.raises(RuntimeError, match="Singular"):
        sm.Canonical(A2, b2, c2, [2, 3], minimize=False).GetBasicSolution()  # duplicate columns
