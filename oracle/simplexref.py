"""ctypes loader for oracle/_ref/libsimplexref.so — the REFERENCE'S OWN sources (problem types,
parser, simplex solver) compiled unmodified behind ref_driver.cpp, with oracle/eigen_shim standing in
for Eigen (which is not on the box).  TEST INFRASTRUCTURE ONLY: importable from tests/ and
__graft_entry__ — never from the product package.

The library is built in the development container, where /root/reference exists (``make -C oracle
ref``); it is git-ignored but travels to the GPU box with the snapshot.  ``available()`` says whether
it can be used here.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libsimplexref.so")
DROPIN = os.path.join(_HERE, "_ref", "dropin_demo")     # INTEGRATION.md's drop-in built with the reference's own types
REFERENCE = os.environ.get("SIMPLEX_REFERENCE", "/root/reference")
COMMON, SYMMETRICAL, CANONICAL = 0, 1, 2
TO_SYMMETRICAL, TO_CANONICAL, TO_COMMON, GET_DUAL = 0, 1, 2, 3
_LIB = None


class Problem(C.Structure):
    _fields_ = [("kind", C.c_int32), ("m", C.c_int32), ("n", C.c_int32), ("maximize", C.c_int32), ("n_orig", C.c_int32),
                ("A", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p),
                ("row_types", C.c_void_p), ("var_types", C.c_void_p), ("basis", C.c_void_p),
                ("cap_m", C.c_int32), ("cap_n", C.c_int32)]


class EnumResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("m", C.c_int32), ("basis", C.c_int32 * 16), ("x_B", C.c_double * 16),
                ("objective", C.c_double), ("best_rank", C.c_uint64), ("n_bases", C.c_uint64),
                ("n_singular", C.c_uint64), ("n_infeasible", C.c_uint64), ("n_feasible", C.c_uint64)]


def build():
    """(Re)build from the reference's sources when they are present; otherwise keep what is there."""
    if os.path.isdir(os.path.join(REFERENCE, "src")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REF={REFERENCE}"])
        if os.path.exists(os.path.join(os.path.dirname(_HERE), "simplexmethod_b200", "libenumgpu.so")):
            subprocess.check_call(["make", "-C", _HERE, "-s", "dropin", f"REF={REFERENCE}"])
    return SO if os.path.exists(SO) else None


def available() -> bool:
    return build() is not None


def linear_algebra() -> str:
    """'shim' (oracle/eigen_shim, the API stand-in) or 'eigen:<include dir>' — what the library was compiled against."""
    try:
        return open(os.path.join(_HERE, "_ref", "EIGEN")).read().strip() or "shim"
    except OSError:
        return "shim"


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libsimplexref.so is absent and the reference sources are not here to build it")
        L = C.CDLL(SO)
        L.ref_last_error.restype = C.c_char_p
        L.ref_convert.argtypes = [C.POINTER(Problem), C.c_int32, C.POINTER(Problem)]
        L.ref_parse.argtypes = [C.c_char_p, C.POINTER(Problem)]
        L.ref_print.argtypes = [C.POINTER(Problem), C.c_char_p, C.c_int32]
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.ref_basic_solution.argtypes = [C.c_int32, C.c_int32, dp, dp, dp, ip, C.c_int32, dp, ip, dp]
        L.ref_enumerate.argtypes = [C.c_int32, C.c_int32, dp, dp, dp, C.c_int32, C.POINTER(C.c_uint8), C.POINTER(EnumResult)]
        L.ref_enumerate_range.argtypes = [C.c_int32, C.c_int32, dp, dp, dp, C.c_int32, C.c_uint64, C.c_uint64,
                                          C.POINTER(C.c_uint8), C.POINTER(EnumResult)]
        L.ref_simplex_solve.argtypes = [C.c_int32, C.c_int32, dp, dp, dp, ip, C.c_int32, C.c_int32, dp]
        _LIB = L
    return _LIB


def last_error() -> str:
    return lib().ref_last_error().decode("utf-8", "replace")


class _Buf:
    """numpy buffers + the struct that points at them."""

    def __init__(self, cap_m, cap_n):
        self.A = np.zeros(cap_m * cap_n)
        self.b = np.zeros(cap_m)
        self.c = np.zeros(cap_n)
        self.rt = np.zeros(cap_m, dtype=np.int32)
        self.vt = np.zeros(cap_n, dtype=np.int32)
        self.basis = np.zeros(cap_m, dtype=np.int32)
        self.s = Problem(0, 0, 0, 0, 0, self.A.ctypes.data, self.b.ctypes.data, self.c.ctypes.data,
                         self.rt.ctypes.data, self.vt.ctypes.data, self.basis.ctypes.data, cap_m, cap_n)

    def fill(self, kind, A, b, c, maximize, row_types=None, var_types=None, basis=None, n_orig=None):
        A = np.asarray(A, dtype=np.float64)
        m, n = A.shape
        self.A[:m * n] = np.asfortranarray(A).ravel(order="F")
        self.b[:m] = b
        self.c[:n] = c
        if row_types is not None:
            self.rt[:m] = row_types
        if var_types is not None:
            self.vt[:n] = var_types
        if basis is not None:
            self.basis[:m] = basis
        self.s.kind, self.s.m, self.s.n, self.s.maximize = kind, m, n, int(bool(maximize))
        self.s.n_orig = n if n_orig is None else n_orig
        return self

    def take(self):
        m, n = self.s.m, self.s.n
        out = {"kind": self.s.kind, "A": self.A[:m * n].reshape((m, n), order="F").copy(), "b": self.b[:m].copy(),
               "c": self.c[:n].copy(), "maximize": bool(self.s.maximize)}
        if self.s.kind == COMMON:
            out["row_types"], out["var_types"] = self.rt[:m].tolist(), self.vt[:n].tolist()
        if self.s.kind == CANONICAL:
            out["basis"], out["n_orig"] = self.basis[:m].tolist(), self.s.n_orig
        return out


def convert(kind, op, A, b, c, maximize, **kw):
    """Run one of the reference's conversions; returns a dict (see _Buf.take) or raises RuntimeError(reference message)."""
    A = np.asarray(A, dtype=np.float64)
    m, n = A.shape
    src = _Buf(m, n).fill(kind, A, b, c, maximize, **kw)
    dst = _Buf(2 * m + 2 * n + 2, 2 * n + 4 * m + 2)
    rc = lib().ref_convert(C.byref(src.s), op, C.byref(dst.s))
    if rc != 0:
        raise RuntimeError(last_error())
    return dst.take()


def print_problem(kind, A, b, c, maximize, **kw) -> str:
    """What the reference's Print() writes to std::cout for this problem."""
    A = np.asarray(A, dtype=np.float64)
    src = _Buf(*A.shape).fill(kind, A, b, c, maximize, **kw)
    buf = C.create_string_buffer(1 << 16)
    n = lib().ref_print(C.byref(src.s), buf, len(buf))
    if n < 0:
        raise RuntimeError(last_error())
    return buf.raw[:n].decode("utf-8")


def parse(text: str):
    """SymmetricalParser::ParseFromString: dict, or None (+ last_error()) as the reference returns nullptr."""
    dst = _Buf(64, 64)
    rc = lib().ref_parse(text.encode("utf-8"), C.byref(dst.s))
    if rc == 1:
        return None
    if rc != 0:
        raise RuntimeError(last_error())
    return dst.take()


def _ptrs(A, b, c):
    A = np.asfortranarray(np.asarray(A, dtype=np.float64))
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    return (A, b, c), (A.ctypes.data_as(dp), b.ctypes.data_as(dp), c.ctypes.data_as(dp))


def basic_solution(A, b, c, basis, minimize=True):
    """Canonical(A,b,c,basis).GetBasicSolution(), IsFeasibleBasis(), Evaluate(x) -> (x, feasible, z)."""
    keep, (pA, pb, pc) = _ptrs(A, b, c)
    m, n = keep[0].shape
    bs = (C.c_int32 * m)(*basis)
    x = np.zeros(n)
    feas, z = C.c_int32(), C.c_double()
    rc = lib().ref_basic_solution(m, n, pA, pb, pc, bs, int(bool(minimize)), x.ctypes.data_as(C.POINTER(C.c_double)),
                                  C.byref(feas), C.byref(z))
    if rc != 0:
        raise RuntimeError(last_error())
    return x, bool(feas.value), z.value


def enumerate_bases(A, b, c, maximize, want_status=False, rank_begin=0, rank_end=0):
    """The enumeration path composed from the reference's per-basis primitives (see ref_driver.cpp);
    ranks [rank_begin, rank_end) of the lexicographic order, (0, 0) = all."""
    from math import comb
    keep, (pA, pb, pc) = _ptrs(A, b, c)
    m, n = keep[0].shape
    count = (rank_end - rank_begin) if (rank_begin or rank_end) else comb(n, m)
    status = np.zeros(max(count, 1), dtype=np.uint8) if want_status else None
    res = EnumResult()
    rc = lib().ref_enumerate_range(m, n, pA, pb, pc, int(bool(maximize)), rank_begin, rank_end,
                                   status.ctypes.data_as(C.POINTER(C.c_uint8)) if want_status else None, C.byref(res))
    if rc != 0:
        raise RuntimeError(last_error())
    return res, status


def simplex_solve(A, b, c, basis, minimize=True, n_orig=None):
    """Solver(Canonical(...)).solve() -> first n_orig components of the optimal x (raises RuntimeError(message))."""
    keep, (pA, pb, pc) = _ptrs(A, b, c)
    m, n = keep[0].shape
    n_orig = n if n_orig is None else n_orig
    bs = (C.c_int32 * m)(*basis)
    x = np.zeros(n_orig)
    rc = lib().ref_simplex_solve(m, n, pA, pb, pc, bs, int(bool(minimize)), n_orig, x.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise RuntimeError(last_error())
    return x
