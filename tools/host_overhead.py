"""Host cost of one compress call: eager C-ABI path vs a captured graph replay (us per call, enqueue only),
and the device time per step of both when issued back to back."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvpress_b200 import native
import bench

native.load()
for wl in os.environ.get("HO_WORKLOADS", "decoding_knorm,knorm_128k,ea_128k,snapkv_32k").split(","):
    w = bench.WORKLOADS[wl]
    K, V, extra = bench.make_inputs(w, "cuda:0", 1)
    n_kept = w.get("n_kept") or bench.kept_count(w["S"], w["ratio"])
    fn = lambda: bench.run_native(w, K, V, extra, n_kept)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    N = 200

    def timed(call):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.record()
        for _ in range(N):
            call()
        e.record()
        host = (time.perf_counter() - t0) / N * 1e6
        torch.cuda.synchronize()
        return host, s.elapsed_time(e) / N * 1e3

    h_e, d_e = timed(fn)
    g = native.capture(fn)
    g.replay()
    h_g, d_g = timed(g.replay)
    print(f"{wl:16s} eager: host {h_e:7.1f} us/call, device {d_e:7.1f} us/step | graph: host {h_g:6.1f} us/call, "
          f"device {d_g:7.1f} us/step", flush=True)
