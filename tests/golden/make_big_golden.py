"""One-off: full CPU-oracle enumeration of the headline LP dense_lp(12, 40, seed=1)
(5 586 853 480 bases, ~10 min on 8 cores) -> tests/golden/dense_12_40_seed1.json.
The GPU parity test at full size compares counters, best rank, basis, x_B and
objective against this file (the oracle itself cannot be re-run inside a test).
The ranks of the singular bases (9 under the default absolute pivot rule) are
recorded too: they are where the absolute rule and the reference's
FullPivLU::isInvertible (SimplexSolover.h:124-126) can disagree, and
tests/test_gpu_parity.py / tests/test_reference_code.py evaluate exactly those.

    python tests/golden/make_big_golden.py [threads]
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import enumcpu  # noqa: E402
from simplexmethod_b200 import lpgen  # noqa: E402


def main():
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else os.cpu_count()
    m, n, seed = 12, 40, 1
    A, b, c, mx = lpgen.dense_lp(m, n, seed)
    t = time.time()
    r, sing, n_sing = enumcpu.list_class(A, b, c, mx, enumcpu.SINGULAR, capacity=4096, n_threads=threads)
    dt = time.time() - t
    assert n_sing == r.n_singular == sing.size
    out = dict(m=m, n=n, seed=seed, status=r.status, n_bases=r.n_bases, n_singular=r.n_singular,
               n_infeasible=r.n_infeasible, n_feasible=r.n_feasible, best_rank=r.best_rank,
               basis=list(r.basis)[:m], x_B=[float(v).hex() for v in list(r.x_B)[:m]],
               objective=float(r.objective).hex(), objective_float=r.objective,
               singular_ranks=[int(v) for v in sing],
               oracle_seconds=dt, oracle_threads=threads)
    with open(os.path.join(HERE, "dense_12_40_seed1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main()
