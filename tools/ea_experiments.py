"""Where does ea_logits_kernel lose its time? Times kvp_expected_attention_score (memset + logits [+ ||v|| side kernel]
+ finalize + sentinel) at the ea_128k shape for the default library and the TIMING-ONLY variants built with
-DKVP_EA_EXP=<bits> (bit 0: epilogue without the k-row LDS, bit 1: no bias MMA step, bit 2: epilogue without
accumulator reads), with and without the concurrent value-norm kernel. Graph replay, us per call."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from kvpress_b200 import native
import bench
native.load()
w = bench.WORKLOADS["ea_128k"]
K, V, extra = bench.make_inputs(w, "cuda:0", 1)
out = []
for vn in (False, True):
    g = native.capture(lambda: native.expected_attention_score(K, V, extra["mu"], extra["cov"], 0.0, 4, vn))
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(40): g.replay()
        e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / 40 * 1e3)
    out.append("vnorm=%%d %%.1f" %% (vn, best))
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
