"""One library, one shape, a few solves through a handle (for ncu):  krun.py lib.so m n [reps] [shard_i shard_n]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from simplexmethod_b200 import _abi, lpgen
L = _abi.bind(C.CDLL(os.path.abspath(sys.argv[1])))
m, n = int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
si, sn = (int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else (0, 0)
A, b, c, mx = lpgen.dense_lp(m, n, 1)
h = C.c_void_p()
assert L.enumgpu_create(-1, C.byref(h)) == 0
ps = _abi.Problem(m, n, m, int(mx), A.ctypes.data, b.ctypes.data, c.ctypes.data)
o = _abi.Options(-1, -1, 0, 0, 0, 0, None, None, si, sn)
for i in range(reps):
    r = _abi.Result()
    assert L.enumgpu_solve_h(h, C.byref(ps), C.byref(o), C.byref(r)) in (0, 1)
    print(f"call {i}: kernel_ms {r.kernel_ms:.3f} bases {r.n_bases} best_rank {r.best_rank} feasible {r.n_feasible}")
L.enumgpu_destroy(h)
