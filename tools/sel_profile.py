"""Phase-level cycle counters of select_compact_kernel (block 0), built with -DKVP_SEL_PROFILE."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["KVPRESS_B200_LIB"] = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "libkvp_prof.so")
import torch
from kvpress_b200 import native
import bench
lib = native.load()
for wl in ("ea_128k", "knorm_128k"):
    w = bench.WORKLOADS[wl]
    K, V, extra = bench.make_inputs(w, "cuda:0", 1)
    n_kept = bench.kept_count(w["S"], w["ratio"])
    for _ in range(3):
        bench.run_native(w, K, V, extra, n_kept)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    lib.kvp_debug_sel_profile(buf, 1)
    bench.run_native(w, K, V, extra, n_kept)
    torch.cuda.synchronize()
    lib.kvp_debug_sel_profile(buf, 0)
    n_c, n_r = max(buf[3], 1), max(buf[6], 1)
    print(f"{wl}: block 0 did {buf[3]} compact items, {buf[6]} refine items; kernel {buf[7]} cycles")
    print(f"  per compact item: wait-ready {buf[0] / n_c:.0f}  rank {buf[1] / n_c:.0f}  copy {buf[2] / n_c:.0f} cycles")
    print(f"  ticket+barriers per pop {buf[4] / (n_c + n_r + 1):.0f}; refine item {buf[5] / n_r:.0f} cycles")
