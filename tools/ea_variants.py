"""Times kvp_expected_attention_score variants at the 128k workload (CUDA events, best of N)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    os.environ["KVPRESS_B200_LIB"] = sys.argv[1]
import torch
from kvpress_b200 import native
import bench
native.load()
w = bench.WORKLOADS["ea_128k"]
K, V, extra = bench.make_inputs(w, "cuda:0", 1)
K2, V2, _ = bench.make_inputs(w, "cuda:0", 2)
def timeit(fn, n=10):
    for _ in range(3): fn(K, V)
    torch.cuda.synchronize()
    ts = []
    for i in range(n):
        a, b = (K, V) if i % 2 == 0 else (K2, V2)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(a, b); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); return round(ts[0], 1), round(ts[len(ts) // 2], 1)
print(os.environ.get("KVPRESS_B200_LIB", "default lib"))
print("score, cov, vnorm   :", timeit(lambda k, v: native.expected_attention_score(k, v, extra["mu"], extra["cov"], 0.0, 4, True)))
print("score, cov, no vnorm:", timeit(lambda k, v: native.expected_attention_score(k, v, extra["mu"], extra["cov"], 0.0, 4, False)))
