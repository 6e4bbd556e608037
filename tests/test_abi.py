"""CPU: the C-ABI library loads, exports every symbol include/enumgpu.h declares,
its struct layouts match the ctypes mirror, its host-only helpers are right, and
— with no GPU — every solve entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import simplexmethod_b200 as sm
from simplexmethod_b200 import _abi, lpgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "enumgpu.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(enumgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    L = sm.lib()
    names = declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(L, name), f"{name} declared in enumgpu.h but not exported"
        assert name in _abi.SYMBOLS, f"{name} has no ctypes prototype"
    assert sorted(_abi.SYMBOLS) == names


def test_struct_layouts_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "enumgpu.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu\\n",sizeof(enumgpu_problem),sizeof(enumgpu_options),'
                   'sizeof(enumgpu_result),sizeof(enumgpu_partial),offsetof(enumgpu_result,best_rank),'
                   'offsetof(enumgpu_partial,basis));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_abi.Problem), C.sizeof(_abi.Options), C.sizeof(_abi.Result), C.sizeof(_abi.Partial),
            _abi.Result.best_rank.offset, _abi.Partial.basis.offset]
    assert got == want


def test_version_and_binomials(oracle):
    L = sm.lib()
    assert L.enumgpu_version() == 200
    assert L.enumgpu_binomial(5, 2) == 10
    assert L.enumgpu_binomial(24, 8) == 735471
    assert L.enumgpu_binomial(30, 10) == 30045015
    assert L.enumgpu_binomial(40, 12) == 5586853480
    assert L.enumgpu_binomial(3, 5) == 0 and L.enumgpu_binomial(65, 2) == 0
    for n in range(0, 65, 7):
        for k in range(0, 17, 3):
            assert L.enumgpu_binomial(n, k) == oracle.lib().enumcpu_binomial(n, k)


def test_rank_unrank_against_oracle(oracle):
    L, O = sm.lib(), oracle.lib()
    rng = np.random.default_rng(0)
    for (n, m) in [(5, 2), (24, 8), (40, 12), (64, 16), (9, 9)]:
        total = L.enumgpu_binomial(n, m)
        S1, S2 = (C.c_int32 * m)(), (C.c_int32 * m)()
        for r in [0, total - 1] + [int(v) for v in rng.integers(0, total, 20)]:
            assert L.enumgpu_unrank(n, m, r, S1) == 0 and O.enumcpu_unrank(n, m, r, S2) == 0
            assert list(S1) == list(S2)
            assert L.enumgpu_rank(n, m, S1) == r
        assert L.enumgpu_unrank(n, m, total, S1) == _abi.ERR_RANGE
    bad = (C.c_int32 * 3)(2, 1, 4)
    assert L.enumgpu_rank(6, 3, bad) == _abi.UINT64_MAX


def test_merge_partial_is_lexicographic_min_and_sum():
    L = sm.lib()
    a, b = _abi.Partial(), _abi.Partial()
    a.key, a.best_rank, a.n_bases, a.n_feasible, a.objective = 1.0, 7, 10, 3, 1.0
    b.key, b.best_rank, b.n_bases, b.n_feasible, b.objective = 1.0, 5, 20, 4, 1.0
    b.basis[0] = 9
    L.enumgpu_merge_partial(C.byref(a), C.byref(b))
    assert (a.key, a.best_rank, a.n_bases, a.n_feasible, a.basis[0]) == (1.0, 5, 30, 7, 9)
    c = _abi.Partial(); c.key, c.best_rank, c.n_bases = float("inf"), _abi.UINT64_MAX, 5
    L.enumgpu_merge_partial(C.byref(a), C.byref(c))
    assert (a.best_rank, a.n_bases) == (5, 35)
    r = _abi.Result()
    L.enumgpu_partial_to_result(C.byref(c), C.byref(r))
    assert r.status == _abi.NO_FEASIBLE
    L.enumgpu_partial_to_result(C.byref(a), C.byref(r))
    assert r.status == _abi.OK and r.best_rank == 5


def test_argument_errors_come_before_cuda():
    """Bad arguments are reported as such whether or not a GPU is present."""
    A, b, c, mx = lpgen.dense_lp(3, 6, 1)
    can = sm.Canonical(A, b, c, [0, 1, 2], minimize=True)
    with pytest.raises(ValueError):
        sm.EnumerationSolver(can).solve(rank_begin=10, rank_end=2)
    bad = A.copy(); bad[0, 0] = np.inf
    with pytest.raises(ValueError):
        sm.EnumerationSolver(sm.Canonical(bad, b, c, [0, 1, 2])).solve()
    with pytest.raises(ValueError):                       # m > n
        sm.EnumerationSolver(sm.Canonical(np.ones((3, 2)), np.ones(3), np.ones(2), [0, 1, 1]))
    with pytest.raises(ValueError):                       # Canonical.cpp:27-46 checks
        sm.Canonical(A, b[:2], c, [0, 1, 2])
    with pytest.raises(ValueError):
        sm.Canonical(A, b, c, [0, 1, 6])


@pytest.mark.skipif(sm.lib().enumgpu_device_count() > 0, reason="a GPU is present")
def test_no_gpu_means_loud_failure_not_fallback():
    A, b, c, mx = lpgen.lab_symmetric_canonical()
    can = sm.Canonical(A, b, c, [3, 4], minimize=False)
    with pytest.raises(sm.EnumGpuError, match="no CPU fallback"):
        sm.EnumerationSolver(can).solve()
    assert sm.lib().enumgpu_fp64_peak_tflops(1) < 0


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "simplexmethod_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("libenumcpu", "enumcpu.h", "enumcpu_", "from oracle", "import oracle", "oracle/_ref"):
                    assert needle not in text, f"{f} references the oracle ({needle})"


def test_work_window_arithmetic_host():
    """k_shared's cost-weighted work windows: weight_unrank inverts weight_of_child, intervals are contiguous
    (simplexmethod_b200/csrc/tests/test_weights.cu, host-only build of the functions the device walks)."""
    csrc = os.path.join(ROOT, "simplexmethod_b200", "csrc")
    subprocess.check_call(["make", "-C", csrc, "-s", "tests/test_weights"])
    out = subprocess.run([os.path.join(csrc, "tests", "test_weights")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "all passed" in out.stdout
