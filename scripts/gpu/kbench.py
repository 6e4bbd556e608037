"""Kernel experiment harness (development aid, run on the GPU box):
    python scripts/gpu/kbench.py scripts/gpu/variants/a.so [b.so ...]
For every library: the headline enumeration (m=12, n=40) — device time of the best of 3 solves through a handle,
checked against the committed golden —, a 1/8 shard of it, and configs 3 (m=10, n=30), 2 (m=8, n=24) and the
degenerate config 5, each checked against the CPU oracle.  One line per library."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from simplexmethod_b200 import _abi, lpgen  # noqa: E402
from oracle import enumcpu  # noqa: E402

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "dense_12_40_seed1.json")))


def run(L, h, lp, reps, opt=None):
    A, b, c, mx = lp
    m, n = A.shape
    ps = _abi.Problem(m, n, m, int(mx), A.ctypes.data, b.ctypes.data, c.ctypes.data)
    best, res = 1e30, None
    for _ in range(reps):
        res = _abi.Result()
        rc = L.enumgpu_solve_h(h, C.byref(ps), C.byref(opt) if opt else None, C.byref(res))
        assert rc in (0, 1), L.enumgpu_last_error()
        best = min(best, res.kernel_ms)
    return best, res


def main():
    small = {k: lp for k, lp in (("10_30", lpgen.dense_lp(10, 30, 1)), ("8_24", lpgen.dense_lp(8, 24, 1)), ("deg", lpgen.degenerate_lp()),
                                 ("7_18", lpgen.small_degenerate_lp()))}
    want = {k: enumcpu.solve(*lp, n_threads=os.cpu_count())[0] for k, lp in small.items()}
    head = lpgen.dense_lp(12, 40, 1)
    for path in sys.argv[1:]:
        L = _abi.bind(C.CDLL(os.path.abspath(path)))
        h = C.c_void_p()
        assert L.enumgpu_create(-1, C.byref(h)) == 0, L.enumgpu_last_error()
        t0 = time.time()
        ms, r = run(L, h, head, 3)
        ok = (r.best_rank == GOLD["best_rank"] and r.n_feasible == GOLD["n_feasible"] and r.n_singular == GOLD["n_singular"]
              and r.n_infeasible == GOLD["n_infeasible"] and float(r.objective).hex() == GOLD["objective"]
              and [float(v).hex() for v in list(r.x_B)[:12]] == GOLD["x_B"])
        o8 = _abi.Options(-1, -1, 0, 0, 0, 0, None, None, 3, 8)
        ms8, _ = run(L, h, head, 3, o8)
        line = f"{os.path.basename(path):28s} 12x40 {ms:8.3f} ms {'OK ' if ok else 'BAD'} algo={r.algo_used}  1/8 {ms8:7.3f}"
        for k, lp in small.items():
            msk, rk = run(L, h, lp, 10)
            w = want[k]
            okk = (rk.best_rank, rk.n_singular, rk.n_infeasible, rk.n_feasible, rk.objective) == \
                  (w.best_rank, w.n_singular, w.n_infeasible, w.n_feasible, w.objective)
            line += f" | {k} {msk:7.4f} {'OK' if okk else 'BAD'}"
        print(line + f"  ({time.time() - t0:.1f}s)", flush=True)
        L.enumgpu_destroy(h)


if __name__ == "__main__":
    main()
