"""Dev: sweep the work-unit size (ENUMGPU_UNIT_MIN / ENUMGPU_UNIT_SHIFT, DEV build only) for a shape and shard count."""
import os, subprocess, sys
lib, m, n = sys.argv[1], sys.argv[2], sys.argv[3]
envs = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[4:]] or [{}]
for env in envs:
    for shard in ((0, 0), (3, 8)):
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, "scripts/gpu/krun.py", lib, m, n, "6", str(shard[0]), str(shard[1])], env=e, capture_output=True, text=True).stdout
        ms = [float(l.split()[3]) for l in out.splitlines() if l.startswith("call")]
        print(m, n, env, "shard", shard, "min ms %.4f" % min(ms[1:]), flush=True)
