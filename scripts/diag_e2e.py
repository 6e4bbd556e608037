"""Where does the host-buffer call spend its time?  wall vs kernel_ms per call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplexmethod_b200 as sm
from simplexmethod_b200 import lpgen, _abi

m, n = int(sys.argv[1]), int(sys.argv[2])
end = int(sys.argv[3]) if len(sys.argv) > 3 else 0
A, b, c, mx = lpgen.dense_lp(m, n, 1)
can = sm.Canonical(A, b, c, list(range(m)), minimize=not mx)
for algo in (_abi.ALGO_SHARED, _abi.ALGO_INDEPENDENT):
    s = sm.EnumerationSolver(can, algo=algo)
    for i in range(4):
        t = time.perf_counter()
        r = s.enumerate(0, end)
        dt = (time.perf_counter() - t) * 1e3
        print(f"algo={algo} call {i}: wall {dt:.3f} ms  kernel_ms {r.kernel_ms:.3f}  launches {r.n_launches} bases {r.n_bases}")
