"""Diagnostic: when does each warp of k_shared start and stop?  Needs scripts/micro/libenumgpu_trace.so
(nvcc ... -DENUMGPU_TRACE, see the Makefile target `trace` in simplexmethod_b200/csrc).
usage: trace_tail.py m n shard_index shard_count   (prints a summary of the LAST launch)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 5 and sys.argv[5] == "child":
    sys.path.insert(0, ROOT)
    from simplexmethod_b200 import _lib
    _lib.LIB_PATH = os.path.join(ROOT, "scripts", "micro", "libenumgpu_trace.so")
    import simplexmethod_b200 as sm
    from simplexmethod_b200 import lpgen
    m, n, si, sc = (int(v) for v in sys.argv[1:5])
    A, b, c, mx = lpgen.dense_lp(m, n, 1)
    s = sm.EnumerationSolver(sm.Canonical(A, b, c, list(range(m)), minimize=not mx), algo=2)
    import ctypes as C
    for i in range(3):
        r = s.enumerate(0, 0, shard_index=si, shard_count=sc)
    print("KERNEL_MS", r.kernel_ms, flush=True)
    buf = (C.c_ulonglong * (4 * 16 * 148))()
    L = C.CDLL(_lib.LIB_PATH)
    assert L.enumgpu_trace_read(buf, len(buf)) == 0
    done = (C.c_ulonglong * 2)()
    L.enumgpu_trace_done(done)
    print("DONE", done[1])
    ph = (C.c_ulonglong * (8 * 16 * 148))()
    L.enumgpu_trace_phase(ph, len(ph))
    tot = [sum(ph[8 * w + i] for w in range(16 * 148) if buf[4 * w + 1]) for i in range(8)]
    print("PHASE", *tot)
    for w in range(16 * 148):
        if buf[4 * w + 1]:
            print("T", w // 16, w % 16, buf[4 * w], buf[4 * w + 1], buf[4 * w + 2], buf[4 * w + 3])
    sys.exit(0)
out = subprocess.run([sys.executable, __file__] + sys.argv[1:5] + ["child"], capture_output=True, text=True).stdout
rows, ms, done, phase = [], None, 0, None
for line in out.splitlines():
    if line.startswith("LAUNCH"):
        rows = []
    elif line.startswith("T "):
        rows.append([int(v) for v in line.split()[1:]])
    elif line.startswith("KERNEL_MS"):
        ms = float(line.split()[1])
    elif line.startswith("DONE"):
        done = int(line.split()[1])
    elif line.startswith("PHASE"):
        phase = [int(v) for v in line.split()[1:]]
import numpy as np
a = np.array(rows, dtype=np.int64)
t0, t1, units, entry = a[:, 2], a[:, 3], a[:, 4], a[:, 5]
g0 = entry.min()
print(f"kernel entry (first CTA) = 0;  CTA entries spread {(entry.max() - g0) / 1e3:.1f} us;  prologue (entry -> unit loop): "
      f"mean {(t0 - entry).mean() / 1e3:.1f} us max {(t0 - entry).max() / 1e3:.1f} us;  record written {(done - g0) / 1e3:.1f} us "
      f"({(done - t1.max()) / 1e3:.1f} us after the last warp stopped)")
end = (t1 - g0) / 1e6
start = (t0 - g0) / 1e6
print(f"warps {len(a)}  kernel_ms(event) {ms:.3f}  span first-start..last-end {end.max():.3f} ms")
print(f"start: max {start.max():.3f} ms   end: min {end.min():.3f} mean {end.mean():.3f} max {end.max():.3f}  -> idle tail mean {end.max() - end.mean():.3f} ms")
print("end percentiles (ms):", " ".join(f"p{p}={np.percentile(end, p):.3f}" for p in (1, 10, 50, 90, 99)))
print(f"units per warp: min {units.min()} mean {units.mean():.1f} max {units.max()}")
if phase and sum(phase):
    names = ["unit fetch + descent", "level q-1 from A", "levels q, q+1", "child / tail-group build", "leaves (setup + d loop + drains)", "end of parent (flush, successor)"]
    tot = sum(phase)
    n_units = int(units.sum())
    print("warp cycles by phase: " + "; ".join(f"{nm} {100 * p / tot:.1f} %" for nm, p in zip(names, phase)))
    print(f"per unit start ({n_units} units): fetch + descent {phase[0] / n_units / 1.965e3:.2f} us")
