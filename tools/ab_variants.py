"""A/B timing of library variants on the same box: each variant runs in a subprocess (its own .so),
bench-style back-to-back steps with CUDA events; reports us/step for ea_128k and knorm_128k."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from kvpress_b200 import native
import bench
native.load()
out = []
for wl in os.environ.get("AB_WORKLOADS", "ea_128k,knorm_128k").split(","):
    w = bench.WORKLOADS[wl]
    K, V, extra = bench.make_inputs(w, "cuda:0", 1)
    n_kept = w.get("n_kept") or bench.kept_count(w["S"], w["ratio"])
    g = native.capture(lambda: bench.run_native(w, K, V, extra, n_kept))   # graph replay: no host launch noise
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
libs = ["default"] + sorted(f for f in os.listdir(os.path.join(ROOT, "tools", "bin")) if f.startswith("libv_"))
for rnd in range(2):
    for lib in libs:
        env = dict(os.environ)
        if lib != "default":
            env["KVPRESS_B200_LIB"] = os.path.join(ROOT, "tools", "bin", lib)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print(f"{lib:22s} {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)

# piggy-back: further time-boxed measurements of the same GPU call (tools/r02_extra.sh), if present
extra = os.path.join(ROOT, "tools", "r02_extra.sh")
if os.environ.get("AB_EXTRA", "1") == "1" and os.path.exists(extra):
    try:
        subprocess.run(["bash", extra], timeout=480)
    except subprocess.TimeoutExpired:
        print("extra script timed out", flush=True)
