"""Small driver for ncu: warm-up call + one measured call of one kernel family.
usage: prof_run.py m n rank_begin rank_end algo(1|2) [repeat] [shard_index shard_count]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplexmethod_b200 as sm
from simplexmethod_b200 import lpgen

m, n, lo, hi, algo = (int(v) for v in sys.argv[1:6])
rep = int(sys.argv[6]) if len(sys.argv) > 6 else 2
shard_i = int(sys.argv[7]) if len(sys.argv) > 7 else 0
shard_n = int(sys.argv[8]) if len(sys.argv) > 8 else 0
A, b, c, mx = lpgen.dense_lp(m, n, 1)
s = sm.EnumerationSolver(sm.Canonical(A, b, c, list(range(m)), minimize=not mx), algo=algo)
for i in range(rep):
    r = s.enumerate(lo, hi, shard_index=shard_i, shard_count=shard_n)
    print(f"call {i}: kernel_ms {r.kernel_ms:.3f} bases {r.n_bases} -> {r.n_bases / r.kernel_ms / 1e6:.2f} G bases/s  best_rank {r.best_rank}")
