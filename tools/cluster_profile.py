"""Phase timeline (globaltimer, ns) of CTA (0,0) of knorm_cluster_kernel at the decoding shape; -DKVP_CL_PROFILE build."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["KVPRESS_B200_LIB"] = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "libv_clprof.so")
import torch
from kvpress_b200 import native
import bench
lib = native.load()
w = bench.WORKLOADS["decoding_knorm"]
K, V, extra = bench.make_inputs(w, "cuda:0", 1)
names = ["entry", "cp.async issued", "cluster.sync #1 done", "loads landed", "scored", "keys pushed", "cluster.sync #2 done",
         "threshold found", "ranked", "stored"]
for rep in range(4):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    bench.run_native(w, K, V, extra, w["n_kept"])
    e.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    lib.kvp_debug_cluster_profile(buf)
    t0 = buf[0]
    print(f"rep {rep}: event time {s.elapsed_time(e) * 1e3:.1f} us; " + ", ".join(f"{n} +{(buf[i] - t0) / 1e3:.2f}" for i, n in enumerate(names)))
