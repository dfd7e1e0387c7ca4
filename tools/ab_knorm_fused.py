"""A/B of the Knorm paths at 128k on one box: two kernels (score, select+compact) vs the fused persistent kernel with
its queue knobs (env, read once per process): lag, head percentage, evict_last on the score-stage K loads."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from kvpress_b200 import native
import bench
native.load()
out = []
for wl in os.environ.get("AB_WORKLOADS", "knorm_128k").split(","):
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
VARIANTS = [
    ("two kernels (default)", {}),
    ("fused lag2", {"KVP_KNORM_FUSED_MAX_MB": "100000", "KVP_KNORM_FUSED_LAG": "2", "KVP_KNORM_FUSED_KLAST": "0"}),
    ("fused lag2 klast", {"KVP_KNORM_FUSED_MAX_MB": "100000", "KVP_KNORM_FUSED_LAG": "2", "KVP_KNORM_FUSED_KLAST": "1"}),
    ("fused lag1 head50", {"KVP_KNORM_FUSED_MAX_MB": "100000", "KVP_KNORM_FUSED_LAG": "1", "KVP_KNORM_FUSED_KLAST": "0"}),
    ("fused lag1 head50 klast", {"KVP_KNORM_FUSED_MAX_MB": "100000", "KVP_KNORM_FUSED_LAG": "1", "KVP_KNORM_FUSED_KLAST": "1"}),
    ("fused lag1 head25 klast", {"KVP_KNORM_FUSED_MAX_MB": "100000", "KVP_KNORM_FUSED_LAG": "1", "KVP_KNORM_FUSED_HEAD": "25", "KVP_KNORM_FUSED_KLAST": "1"}),
    ("fused lag1 head100 klast", {"KVP_KNORM_FUSED_MAX_MB": "100000", "KVP_KNORM_FUSED_LAG": "1", "KVP_KNORM_FUSED_HEAD": "100", "KVP_KNORM_FUSED_KLAST": "1"}),
]
for rnd in range(2):
    for name, env_extra in VARIANTS:
        env = dict(os.environ, **env_extra)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print(f"{name:28s} {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
