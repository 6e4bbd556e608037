"""ctypes loader for the CPU oracle library (TEST INFRASTRUCTURE ONLY).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs — never from the product package.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from simplexmethod_b200 import _abi  # noqa: E402  (struct layouts are shared by design)

FEASIBLE, INFEASIBLE, SINGULAR = 0, 1, 2
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libenumcpu.so")
    src = [os.path.join(_HERE, f) for f in ("enumcpu.c", "enumcpu.h")] + \
          [os.path.join(os.path.dirname(_HERE), "include", "enumgpu.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libenumcpu.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.enumcpu_binomial.restype = C.c_uint64
        L.enumcpu_binomial.argtypes = [C.c_int32, C.c_int32]
        L.enumcpu_rank.restype = C.c_uint64
        L.enumcpu_rank.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        L.enumcpu_unrank.restype = C.c_int
        L.enumcpu_unrank.argtypes = [C.c_int32, C.c_int32, C.c_uint64, C.POINTER(C.c_int32)]
        L.enumcpu_scale.restype = C.c_double
        L.enumcpu_scale.argtypes = [C.POINTER(_abi.Problem)]
        L.enumcpu_eval_basis.restype = C.c_int
        L.enumcpu_eval_basis.argtypes = [C.POINTER(_abi.Problem), C.c_double, C.c_double,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.enumcpu_solve.restype = C.c_int
        L.enumcpu_solve.argtypes = [C.POINTER(_abi.Problem), C.POINTER(_abi.Options), C.POINTER(_abi.Result)]
        L.enumcpu_solve_ex.restype = C.c_int
        L.enumcpu_solve_ex.argtypes = [C.POINTER(_abi.Problem), C.POINTER(_abi.Options), C.c_int,
                                       C.POINTER(C.c_uint8), C.POINTER(_abi.Result)]
        L.enumcpu_eval_basis_rule.restype = C.c_int
        L.enumcpu_eval_basis_rule.argtypes = [C.POINTER(_abi.Problem), C.c_double, C.c_int, C.c_double,
                                              C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.enumcpu_solve_list.restype = C.c_int
        L.enumcpu_solve_list.argtypes = [C.POINTER(_abi.Problem), C.POINTER(_abi.Options), C.c_int, C.POINTER(C.c_uint8),
                                         C.c_int, C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(C.c_uint64),
                                         C.POINTER(_abi.Result)]
        _LIB = L
    return _LIB


class HostProblem:
    """Keeps numpy buffers alive next to the ctypes struct that points at them."""

    def __init__(self, A, b, c, maximize):
        self.A = np.asfortranarray(np.asarray(A, dtype=np.float64))
        self.b = np.ascontiguousarray(np.asarray(b, dtype=np.float64))
        self.c = np.ascontiguousarray(np.asarray(c, dtype=np.float64))
        m, n = self.A.shape
        self.m, self.n = m, n
        self.struct = _abi.Problem(m, n, m, int(bool(maximize)),
                                   self.A.ctypes.data, self.b.ctypes.data, self.c.ctypes.data)


def make_options(eps_feas=-1.0, eps_piv=-1.0, rank_begin=0, rank_end=0, algo=0, pivot_rule=0):
    return _abi.Options(eps_feas, eps_piv, rank_begin, rank_end, 0, algo, None, None, 0, 0, pivot_rule, 0)


def solve(A, b, c, maximize, n_threads=1, want_status=False, **opt):
    """Enumerate with the oracle.  Returns (Result, status_bytes|None)."""
    hp = HostProblem(A, b, c, maximize)
    o = make_options(**opt)
    res = _abi.Result()
    status = None
    ptr = None
    if want_status:
        total = lib().enumcpu_binomial(hp.n, hp.m)
        lo, hi = o.rank_begin, o.rank_end
        if lo == 0 and hi == 0:
            hi = total
        status = np.zeros(max(hi - lo, 1), dtype=np.uint8)
        ptr = status.ctypes.data_as(C.POINTER(C.c_uint8))
    lib().enumcpu_solve_ex(C.byref(hp.struct), C.byref(o), int(n_threads), ptr, C.byref(res))
    return res, status


def list_class(A, b, c, maximize, cls, capacity=1 << 20, n_threads=1, **opt):
    """Enumerate and list the ranks (ascending) of the bases of class `cls` (FEASIBLE / INFEASIBLE / SINGULAR).
    Returns (Result, ranks uint64 array, full count)."""
    hp = HostProblem(A, b, c, maximize)
    o = make_options(**opt)
    res = _abi.Result()
    ranks = np.zeros(max(int(capacity), 1), dtype=np.uint64)
    n_listed = C.c_uint64()
    lib().enumcpu_solve_list(C.byref(hp.struct), C.byref(o), int(n_threads), None, int(cls),
                             ranks.ctypes.data_as(C.POINTER(C.c_uint64)), int(capacity), C.byref(n_listed), C.byref(res))
    return res, ranks[: min(n_listed.value, int(capacity))].copy(), n_listed.value


def eval_basis(A, b, c, maximize, S, eps_feas=1e-9, eps_piv=None, pivot_rule=0):
    hp = HostProblem(A, b, c, maximize)
    if pivot_rule == _abi.PIVOT_RELATIVE:
        tol = hp.m * 2.0 ** -52 if eps_piv is None else eps_piv
    else:
        tol = (1e-9 if eps_piv is None else eps_piv) * lib().enumcpu_scale(C.byref(hp.struct))
    Sv = (C.c_int32 * hp.m)(*S)
    x = (C.c_double * hp.m)()
    z = C.c_double()
    st = lib().enumcpu_eval_basis_rule(C.byref(hp.struct), eps_feas, int(pivot_rule), tol, Sv, x, C.byref(z))
    return st, list(x), z.value
