"""A/B timing of ENV knobs on the same box: python tools/ab_env.py "NAME=VAL[,NAME2=VAL2]" ... ; each setting runs in a
subprocess (knobs are read once per process), graph replay, us/step, best of 3 x 40 steps. AB_WORKLOADS selects workloads."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from kvpress_b200 import native
import bench
native.load()
out = []
for wl in os.environ.get("AB_WORKLOADS", "ea_128k").split(","):
    w = bench.WORKLOADS[wl]
    K, V, extra = bench.make_inputs(w, "cuda:0", 1)
    n_kept = w.get("n_kept") or bench.kept_count(w["S"], w["ratio"])
    g = native.capture(lambda: bench.run_native(w, K, V, extra, n_kept))
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(40): g.replay()
        e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / 40 * 1e3)
    out.append("%%s %%.1f" %% (wl, best))
print("  ".join(out))
''' % ROOT
settings = sys.argv[1:] or ["default"]
for rnd in range(2):
    for st in settings:
        env = dict(os.environ)
        if st != "default":
            for kv in st.split(","):
                k, v = kv.split("=", 1)
                env[k] = v
        try:
            r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=150)
            print(f"{st:28s} {r.stdout.strip() or r.stderr.strip()[-400:]}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"{st:28s} TIMEOUT", flush=True)
