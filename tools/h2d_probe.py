"""Host-link ceiling of the e2e leg at N ranks (torchrun): every rank moves exactly the bytes one ea_128k e2e step moves
(K and V in from pinned host memory, K' and V' out) with NO kernels in between — first rank 0 alone, then all ranks
together. together / alone is the best e2e scaling efficiency the host side of the box allows."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
numa = bench.bind_to_gpu_numa(local)
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
w = bench.WORKLOADS["ea_128k"]
n_kept = bench.kept_count(w["S"], w["ratio"])
shape_in, shape_out = (w["B"], w["Hkv"], w["S"], w["D"]), (w["B"], w["Hkv"], n_kept, w["D"])
Kh = torch.zeros(shape_in, dtype=torch.bfloat16).pin_memory()
Vh = torch.zeros(shape_in, dtype=torch.bfloat16).pin_memory()
Ko = torch.empty(shape_out, dtype=torch.bfloat16).pin_memory()
Vo = torch.empty(shape_out, dtype=torch.bfloat16).pin_memory()
Kd, Vd = torch.empty(shape_in, dtype=torch.bfloat16, device=dev), torch.empty(shape_in, dtype=torch.bfloat16, device=dev)
Kd2, Vd2 = torch.zeros(shape_out, dtype=torch.bfloat16, device=dev), torch.zeros(shape_out, dtype=torch.bfloat16, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

def step():
    with torch.cuda.stream(s_in):
        Kd.copy_(Kh, non_blocking=True); Vd.copy_(Vh, non_blocking=True)
    with torch.cuda.stream(s_out):
        Ko.copy_(Kd2, non_blocking=True); Vo.copy_(Vd2, non_blocking=True)

def timed(active, steps=8):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if active:
        for _ in range(steps): step()
        torch.cuda.current_stream().wait_stream(s_in); torch.cuda.current_stream().wait_stream(s_out)
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier(); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

for _ in range(2): step()
torch.cuda.synchronize()
alone = timed(rank == 0)
together = timed(True)
if rank == 0:
    h2d, d2h = 2 * Kh.numel() * 2, 2 * Ko.numel() * 2
    print(json.dumps({"n_ranks": world, "bytes_h2d_per_step": h2d, "bytes_d2h_per_step": d2h,
                      "rank0_alone_ms": alone, "all_ranks_ms": together, "h2d_GBps_alone": h2d / alone / 1e6,
                      "h2d_GBps_per_rank_together": h2d / together / 1e6,
                      "aggregate_host_GBps_together": world * (h2d + d2h) / together / 1e6,
                      "copy_only_scaling_ceiling": alone / together, "numa_binding": numa}))
if world > 1: dist.destroy_process_group()
