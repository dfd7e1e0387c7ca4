"""Role-level wait/compute cycle counters of the ExpectedAttention logits kernel (block 0), -DKVP_EA_PROFILE build.
KVP_EA_PAIR=0 profiles the one-CTA kernel, the default the CTA-pair kernel."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["KVPRESS_B200_LIB"] = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "libv_prof.so")
import torch
from kvpress_b200 import native
import bench
lib = native.load()
w = bench.WORKLOADS["ea_128k"]
K, V, extra = bench.make_inputs(w, "cuda:0", 1)
n_kept = bench.kept_count(w["S"], w["ratio"])
for _ in range(3):
    bench.run_native(w, K, V, extra, n_kept)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
lib.kvp_debug_ea_profile(buf, 1)
bench.run_native(w, K, V, extra, n_kept)
torch.cuda.synchronize()
lib.kvp_debug_ea_profile(buf, 0)
names = ["producer wait k_empty", "mma wait k_full", "mma wait t_empty", "epi(WG0 w4) wait k_full", "epi(w4) wait t_full",
         "epi(w4) drain compute", "epi(w4) wait v_full", "thread 0: entry -> exit", "thread 128 (epi): entry -> exit",
         "epi warp: entry -> first tile", "epi(w4) v-norm compute", "epi(w4) stores+softmax", "epi w4: entry -> loop end",
         "epi w8: entry -> loop end", "items (tile pairs) of block 0"]
print("pair kernel" if os.environ.get("KVP_EA_PAIR", "1") != "0" else "one-CTA kernel")
for i, n in enumerate(names):
    print(f"{n:32s} {buf[i]:10d}")
