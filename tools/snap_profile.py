"""Role-level wait/compute cycle counters of snap_stats_kernel (block 0), built with -DKVP_SNAP_PROFILE."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["KVPRESS_B200_LIB"] = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "libv_snapprof.so")
import torch
from kvpress_b200 import native
import bench
lib = native.load()
for wl in ("snapkv_32k", "snapkv_128k_70b"):
    w = bench.WORKLOADS[wl]
    K, V, extra = bench.make_inputs(w, "cuda:0", 1)
    n_kept = bench.kept_count(w["S"], w["ratio"])
    for _ in range(3):
        bench.run_native(w, K, V, extra, n_kept)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    lib.kvp_debug_snap_profile(buf, 1)
    bench.run_native(w, K, V, extra, n_kept)
    torch.cuda.synchronize()
    lib.kvp_debug_snap_profile(buf, 0)
    names = ["producer wait k_empty", "mma wait k_full", "mma wait t_empty", "epi(w4) wait t_full", "epi(w4) compute"]
    tiles = -(-(w["S"] // 128) // (148 // 8))
    print(wl, "tiles per CTA ~", tiles)
    for i, n in enumerate(names):
        print(f"  {n:24s} {buf[i]:10d} cycles  = {buf[i] / tiles:9.0f} per tile")
