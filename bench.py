#!/usr/bin/env python
"""bench.py — throughput of the fused score -> top-k -> gather path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one ScorerPress.compress call on one layer's synthetic cache [B, Hkv, S, D] (bf16).
Prints ONE JSON line on rank 0:
  value      tokens scored+compacted per second, whole job (all ranks), inputs resident in HBM
  e2e        the same metric through the public press API with pinned HOST K/V copied in and the
             compacted K'/V' copied back inside the timed region
  roofline   algorithmic bytes of one compress call / its CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (port of the reference's ATen sequence) on the host cores
Multi-GPU: the path has no exchange step; ranks process independent batch shards (weak scaling),
NCCL is used only for the barrier and the max-over-ranks of the timings.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

L2_BYTES = 126 * 1024 * 1024

# name -> workload description. `config_index` refers to BASELINE.json "configs".
WORKLOADS = {
    # BASELINE.json configs[2]: the 128k single-GPU configuration the metric is quoted on
    "ea_128k": dict(scorer="expected_attention", B=1, Hkv=8, Hq=32, S=131072, D=128, ratio=0.7, config_index=2,
                    label="ExpectedAttentionPress r=0.7, Llama-3.1-8B layer shape, 128k ctx"),
    # BASELINE.json configs[1]
    "snapkv_32k": dict(scorer="snapkv", B=1, Hkv=8, Hq=32, S=32768, D=128, ratio=0.5, config_index=1,
                       label="SnapKVPress r=0.5, Llama-3.1-8B layer shape, 32k ctx"),
    # per-layer shape of configs[4] (70B: Hq=64)
    "snapkv_128k_70b": dict(scorer="snapkv", B=1, Hkv=8, Hq=64, S=131072, D=128, ratio=0.5, config_index=4,
                            label="SnapKVPress r=0.5, Llama-3.1-70B layer shape, 128k ctx"),
    "knorm_128k": dict(scorer="knorm", B=1, Hkv=8, Hq=8, S=131072, D=128, ratio=0.5, config_index=None,
                       label="KnormPress r=0.5, Llama-3.1-8B layer shape, 128k ctx"),
    # KeyRerotationPress(KnormPress): score + select + compaction with re-rotated keys
    "rerotate_knorm_128k": dict(scorer="knorm_rerotate", B=1, Hkv=8, Hq=8, S=131072, D=128, ratio=0.5,
                                config_index=None, label="KeyRerotationPress(KnormPress) r=0.5, 128k ctx"),
    # AdaKVPress(ExpectedAttentionPress): EA scores, then the two head-wise selections (no compaction)
    "adakv_ea_128k": dict(scorer="adakv_ea", B=1, Hkv=8, Hq=32, S=131072, D=128, ratio=0.7, config_index=None,
                          label="AdaKVPress(ExpectedAttentionPress) r=0.7 head-wise selection, 128k ctx"),
    # KeyDiffPress (two streaming passes over K, then select + compact)
    "keydiff_128k": dict(scorer="keydiff", B=1, Hkv=8, Hq=8, S=131072, D=128, ratio=0.5, config_index=None,
                         label="KeyDiffPress r=0.5, Llama-3.1-8B layer shape, 128k ctx"),
    "streaming_128k": dict(scorer="streaming", B=1, Hkv=8, Hq=8, S=131072, D=128, ratio=0.5, config_index=None,
                           label="StreamingLLMPress r=0.5, 128k ctx"),
    # steady state of configs[3]: DecodingPress(Knorm, 512, 2048) compaction 2560 -> 2048
    "decoding_knorm": dict(scorer="knorm", B=1, Hkv=8, Hq=8, S=2560, D=128, ratio=None, n_kept=2048,
                           config_index=3, label="DecodingPress(Knorm) steady-state compaction 2560->2048"),
}
DEFAULT_WORKLOAD = "ea_128k"


def kept_count(S: int, ratio: float) -> int:
    return int(S * (1 - ratio))


def algorithmic_bytes(w: dict, n_kept: int) -> int:
    """every input byte read once + every output byte written once, per (b, kv-head)."""
    row = w["D"] * 2
    S = w["S"]
    per_head = {
        "knorm": row * (S + 3 * n_kept),
        "knorm_rerotate": row * (S + 3 * n_kept),
        "keydiff": row * (2 * S + 3 * n_kept),  # the anchor needs all of K before any score: K is read twice
        "adakv_ea": row * 2 * S,  # read all K and all V once; the output is index triples only
        "snapkv": row * (S + 3 * n_kept),
        "expected_attention": row * (2 * S + 2 * n_kept),
        "streaming": row * 4 * n_kept,
    }[w["scorer"]]
    return per_head * w["B"] * w["Hkv"]


def algorithmic_flops(w: dict) -> float:
    """dense FLOPs of the tensor-core score stage (0 for the streaming scorers)."""
    B, Hq, S, D = w["B"], w["Hq"], w["S"], w["D"]
    if w["scorer"] == "expected_attention":
        return float(B * Hq * S * (2 * D * D + 4 * D))
    if w["scorer"] == "snapkv":
        return float(2 * 2 * 64 * B * Hq * S * D)  # two exact-softmax passes of 2*w*Hq*S*D
    return 0.0


def load_traffic(workload: str):
    """DRAM bytes one compress call moves (sum over its kernels, dram__bytes_read + write per launch) from the
    committed ncu launch list of the same command: profiles/traffic.json, written by tools/ncu_launches.py."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        return json.loads(p.read_text()).get(workload)
    return None


def load_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops", 1590.0)),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------------------
def make_inputs(w: dict, device, seed: int, pinned_host: bool = False):
    """K, V ~ N(0,1) bf16 [B,Hkv,S,D] (+ scorer-specific small operands), seeded per rank."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    B, H, Hq, S, D = w["B"], w["Hkv"], w["Hq"], w["S"], w["D"]
    if device != "cpu" and not pinned_host:
        gd = torch.Generator(device=device).manual_seed(seed)
        K = torch.randn((B, H, S, D), generator=gd, device=device, dtype=torch.float32).to(torch.bfloat16)
        V = torch.randn((B, H, S, D), generator=gd, device=device, dtype=torch.float32).to(torch.bfloat16)
    else:
        K = torch.randn((B, H, S, D), generator=g, dtype=torch.float32).to(torch.bfloat16)
        V = torch.randn((B, H, S, D), generator=g, dtype=torch.float32).to(torch.bfloat16)
        if pinned_host:
            K, V = K.pin_memory(), V.pin_memory()
    extra = {}
    if w["scorer"] == "knorm_rerotate":
        extra["inv_freq"] = 1.0 / (500000.0 ** (torch.arange(0, D, 2).float() / D))  # Llama-3 rope_theta
    if w["scorer"] == "snapkv":
        extra["q_window"] = torch.randn((B, Hq, 64, D), generator=g, dtype=torch.float32).to(torch.bfloat16)
    if w["scorer"] in ("expected_attention", "adakv_ea"):
        extra["mu"] = (0.5 * torch.randn((B, Hq, D), generator=g)).to(torch.bfloat16)
        a = torch.randn((B, Hq, D, D), generator=g) / D ** 0.5
        extra["cov"] = (a @ a.transpose(-1, -2)).to(torch.bfloat16)  # PSD like a real covariance
    if device != "cpu":
        extra = {k: v.to(device) for k, v in extra.items()}
    return K, V, extra


def run_native(w: dict, K, V, extra, n_kept: int):
    from kvpress_b200 import native

    s = w["scorer"]
    if s == "knorm":
        return native.knorm_compress(K, V, n_kept)[:2]
    if s == "keydiff":
        return native.keydiff_compress(K, V, n_kept)[:2]
    if s == "knorm_rerotate":
        return native.scores_compress_rerotate(native.knorm_score(K), K, V, n_kept, extra["inv_freq"])[:2]
    if s == "adakv_ea":
        sc = native.expected_attention_score(K, V, extra["mu"], extra["cov"], 1e-2, 4, True)
        n_safe = int(n_kept * 0.2)
        sc = sc.scatter(-1, native.scores_select(sc, n_safe).long(), torch.finfo(sc.dtype).max)
        H, S = sc.shape[1], sc.shape[2]
        return native.scores_select((-sc).reshape(sc.shape[0], 1, H * S), H * (S - n_kept)), None
    if s == "streaming":
        return native.streaming_compress(K, V, n_kept, 4)[:2]
    if s == "snapkv":
        return native.snapkv_compress(K, V, extra["q_window"], 64, 5, n_kept)[:2]
    if s == "expected_attention":
        return native.expected_attention_compress(K, V, extra["mu"], extra["cov"], 0.0, 4, True, n_kept)[:2]
    raise ValueError(s)


def run_oracle(w: dict, K, V, extra, ratio: float):
    from oracle import press_oracle as O

    s = w["scorer"]
    if s == "knorm":
        return O.knorm_compress(K, V, ratio)
    if s == "keydiff":
        n_kept = O.kept_count(K.shape[2], ratio)
        idx = O.topk_indices(O.keydiff_scores(K), n_kept)
        return O.gather_rows(K, idx), O.gather_rows(V, idx)
    if s == "knorm_rerotate":
        n_kept = O.kept_count(K.shape[2], ratio)
        return O.key_rerotation_compress(O.knorm_scores(K), K, V, n_kept, extra["inv_freq"])[:2]
    if s == "adakv_ea":
        sc = O.expected_attention_scores(K, V, extra["mu"], extra["cov"], 1e-2, 4, True)
        n_kept = O.kept_count(K.shape[2], ratio)
        sc = sc.scatter(-1, sc.topk(int(n_kept * 0.2), dim=-1).indices, torch.finfo(sc.dtype).max)
        H, S = sc.shape[1], sc.shape[2]
        return torch.topk(-sc.reshape(sc.shape[0], -1), H * (S - n_kept), dim=1).indices, None
    if s == "streaming":
        return O.streaming_compress(K, V, ratio, 4)
    if s == "snapkv":
        return O.snapkv_compress(extra["q_window"], K, V, ratio, 64, 5)
    if s == "expected_attention":
        return O.expected_attention_compress(K, V, extra["mu"], extra["cov"], ratio)
    raise ValueError(s)


def effective_ratio(w: dict) -> float:
    if w.get("ratio") is not None:
        return w["ratio"]
    from oracle import press_oracle as O

    return O.find_target_compression_ratio(w["S"], w["n_kept"])


# --------------------------------------------------------------------------------------------------
# clocks sampler (pynvml; nvidia-smi columns of B200_PROFILING.md)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown",
    }

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:  # no NVML: report nulls rather than fail the bench
            self._nv = None

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self) -> dict:
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's ATen sequence on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_leg(w: dict, budget_s: float, max_reps: int = 3):
    """Times oracle compress on a bounded sample of the workload (S shrunk until one call fits the
    budget). Returns (tokens_per_s, cores, sample description, seconds per call)."""
    torch.set_num_threads(os.cpu_count() or 1)
    ratio = effective_ratio(w)
    # rough cost model (ms per 1k tokens on ~8 cores) to size the sample without trial runs
    per_k = {"knorm": 3.5, "keydiff": 8.0, "knorm_rerotate": 6.0, "adakv_ea": 50.0, "streaming": 3.0, "snapkv": 16.0, "expected_attention": 50.0}[w["scorer"]]
    S = w["S"]
    while S > 4096 and per_k * S / 1000 / 1000 * max_reps > budget_s:
        S //= 2
    ws = dict(w, S=S)
    K, V, extra = make_inputs(ws, "cpu", 1234)
    best = float("inf")
    run_oracle(ws, K[:, :, : min(S, 2048)], V[:, :, : min(S, 2048)], extra, ratio)  # warm the thread pool
    t_start = time.perf_counter()
    for _ in range(max_reps):
        t0 = time.perf_counter()
        run_oracle(ws, K, V, extra, ratio)
        best = min(best, time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    tokens = ws["B"] * S
    sample = f"{w['scorer']} oracle compress on [B={ws['B']},Hkv={ws['Hkv']},S={S},D={ws['D']}] bf16, best of <= {max_reps}"
    return tokens / best, torch.get_num_threads(), sample, best


# --------------------------------------------------------------------------------------------------
# multi-rank bookkeeping (no data-path collective: ranks own disjoint batch shards)
# --------------------------------------------------------------------------------------------------
def rank_seed(rank: int, i: int = 0) -> int:
    """Every rank synthesises its own shard; seeds never collide across ranks / input sets."""
    return 1234 + 1000 * rank + i


def max_over_ranks(ms_local: float, dist, device) -> float:
    """The step time of the job is the slowest rank's device time."""
    t = torch.tensor([ms_local], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_tokens_per_s(tokens_per_rank_step: int, world: int, ms_per_step: float) -> float:
    """Weak scaling: every rank processes tokens_per_rank_step per step; value is the job aggregate."""
    return tokens_per_rank_step * world / (ms_per_step * 1e-3)


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=os.environ.get("KVP_BENCH_WORKLOAD", DEFAULT_WORKLOAD), choices=list(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-mode", default="auto", choices=["auto", "serial", "staged", "zero_copy"],
                    help="host path: serial = H2D, compress, D2H on one stream; staged = per-kv-head 3-stream pipeline "
                         "(kvpress_b200.host_staging); zero_copy = staged + kept V rows read in place over PCIe")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    ratio = effective_ratio(w)
    n_kept = w.get("n_kept") or kept_count(w["S"], ratio)
    metric = "KV tokens scored+compacted/sec"
    config = {
        "workload": args.workload, "description": w["label"], "baseline_config_index": w["config_index"],
        "B_per_gpu": w["B"], "Hkv": w["Hkv"], "Hq": w["Hq"], "S": w["S"], "D": w["D"], "n_kept": n_kept,
        "compression_ratio": ratio, "sharding": f"batch over {world} rank(s), no data-path collective",
    }

    # ---------------- reference arm: CPU oracle, rank 0 only -------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        value, cores, sample, secs = cpu_leg(w, budget_s=60.0, max_reps=max(args.warmup, 0) + steps)
        line = {
            "impl": "reference", "metric": metric, "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return

    # ---------------- B200 arm ---------------------------------------------------------------------
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device(device))

    from kvpress_b200 import native

    native.load()

    # enough distinct input sets that consecutive steps never find their K/V in L2
    bytes_per_set = 2 * w["B"] * w["Hkv"] * w["S"] * w["D"] * 2
    n_sets = max(1, min(8, -(-4 * L2_BYTES // bytes_per_set)))
    sets = [make_inputs(w, device, rank_seed(rank, i)) for i in range(n_sets)]
    flush_note = (f"{n_sets} rotating input sets x {bytes_per_set / 2**20:.0f} MiB (> L2) so no step re-reads "
                  "cached K/V" if n_sets > 1 else f"inputs {bytes_per_set / 2**20:.0f} MiB per step > 126 MiB L2")
    config["l2"] = flush_note

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        K, V, extra = sets[i % n_sets]
        run_native(w, K, V, extra, n_kept)
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        start.record()
        for i in range(args.steps):
            K, V, extra = sets[i % n_sets]
            run_native(w, K, V, extra, n_kept)
        stop.record()
        torch.cuda.synchronize()
    barrier()
    ms_total = max_over_ranks(start.elapsed_time(stop), dist, device)
    ms_per_step = ms_total / args.steps
    tokens_per_step = w["B"] * w["S"] * world
    value = whole_job_tokens_per_s(w["B"] * w["S"], world, ms_per_step)

    # roofline of one compress call on one GPU
    peaks = load_peaks()
    abytes = algorithmic_bytes(w, n_kept)
    achieved = abytes / (ms_per_step * 1e-3) / 1e9
    p = native.make_problem(sets[0][0], sets[0][1], n_kept, w["Hq"])
    if w["scorer"] == "knorm_rerotate":  # kvp_knorm_score (1) + kvp_scores_compress_rerotate (generic: 3)
        launches = 1 + native.launches_per_compress(p, 0)
    elif w["scorer"] == "adakv_ea":  # EA score (memset, logits, vnorm, finalize, sentinel) + 2 x kvp_scores_select (3 each)
        launches = 5 + 2 * native.launches_per_compress(p, 0)
    else:
        scorer_id = {"knorm": 1, "streaming": 2, "snapkv": 3, "expected_attention": 4, "keydiff": 5}[w["scorer"]]
        launches = native.launches_per_compress(p, scorer_id)
    roofline = {
        "bound": "hbm", "kernel": f"kvp_{w['scorer']}_compress ({launches} launches: score, select, compact)",
        "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
        "peak_source": peaks["source"], "algorithmic_bytes_per_launch": abytes, "traffic": None,
    }
    tr = load_traffic(args.workload)
    if tr is not None:
        roofline["traffic"] = tr["dram_bytes_per_call"]
        roofline["traffic_source"] = tr["source"]
        roofline["kernel_us_ncu"] = tr.get("kernel_us")
    flops = algorithmic_flops(w)
    if flops:
        # lower bound on the tensor-core rate of the score stage: all of the step time charged to it
        roofline["tensor"] = {"algorithmic_flops_per_launch": flops, "achieved_tflops_lower_bound": flops / (ms_per_step * 1e-3) / 1e12,
                              "peak_tflops": peaks["bf16_tflops"], "unit": "TFLOP/s"}

    # ---------------- e2e: pinned host K/V in, compacted K'/V' out, through the press API ---------
    e2e = None
    if not args.no_e2e:
        Kh, Vh, extra_h = make_inputs(w, device, 99 + rank, pinned_host=True)
        out_k = torch.empty((w["B"], w["Hkv"], n_kept, w["D"]), dtype=torch.bfloat16).pin_memory()
        out_v = torch.empty_like(out_k).pin_memory()
        out_idx = torch.empty(w["B"] * w["Hkv"] * (w["S"] - n_kept), dtype=torch.int32).pin_memory()
        e2e_steps = max(2, min(args.steps, 8))

        from kvpress_b200 import host_staging

        mode = args.e2e_mode
        if mode == "auto":
            # measured (profiles/r01_e2e_modes.txt): zero_copy > staged > serial wherever V is not read to score
            mode = {"knorm_rerotate": "serial", "adakv_ea": "serial", "expected_attention": "staged"}.get(w["scorer"], "zero_copy")
        if mode != "serial" and w["scorer"] in ("knorm_rerotate", "adakv_ea"):
            raise SystemExit("this workload only has the serial host path")
        if mode == "zero_copy" and w["scorer"] == "expected_attention":
            raise SystemExit("zero_copy needs a scorer that does not read V to score")

        extra_pinned = {k: v.cpu().pin_memory() for k, v in extra_h.items()}  # small operands travel every step too

        def e2e_step():
            extra_d = {k: v.to(device, non_blocking=True) for k, v in extra_pinned.items()}
            if mode == "serial":
                Kd = Kh.to(device, non_blocking=True)
                Vd = Vh.to(device, non_blocking=True)
                k2, v2 = run_native(w, Kd, Vd, extra_d, n_kept)
                if v2 is None:  # head-wise selection: the result is the pruned index list
                    out_idx.copy_(k2.reshape(-1), non_blocking=True)
                    return
                out_k.copy_(k2, non_blocking=True)
                out_v.copy_(v2, non_blocking=True)
                return
            params = dict(extra_d)
            if w["scorer"] == "snapkv":
                params.update(window=64, kernel_size=5)
            if w["scorer"] == "expected_attention":
                params.update(epsilon=0.0, n_sink=4, use_vnorm=True)
            host_staging.compress_host(w["scorer"], Kh, Vh, n_kept, device=device, out_keys=out_k, out_values=out_v,
                                       heads_per_chunk=1, values_zero_copy=(mode == "zero_copy"), **params)

        for _ in range(3):
            e2e_step()
        barrier()
        start.record()
        for _ in range(e2e_steps):
            e2e_step()
        stop.record()
        torch.cuda.synchronize()
        barrier()
        e2e_ms = max_over_ranks(start.elapsed_time(stop), dist, device) / e2e_steps
        small = sum(v.numel() * v.element_size() for v in extra_pinned.values())
        h2d_bytes = small + {"serial": 2 * Kh.numel() * 2, "staged": 2 * Kh.numel() * 2,
                             "zero_copy": (2 * out_k.numel() * 2 if w["scorer"] == "streaming"
                                           else Kh.numel() * 2 + out_v.numel() * 2)}[mode]
        e2e = {
            "value": tokens_per_step / (e2e_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms, "steps": e2e_steps,
            "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": (out_idx.numel() * 4 if w["scorer"] == "adakv_ea" else 2 * out_k.numel() * 2),
            "mode": mode,
            "path": {"serial": "pinned host K,V -> H2D -> kvp_*_compress -> D2H K',V' on one stream",
                     "staged": "kvpress_b200.host_staging.compress_host: per-kv-head chunks, H2D | kvp_*_compress | D2H "
                               "on three streams",
                     "zero_copy": "host_staging.compress_host: K staged per kv-head, kept V rows gathered in place "
                                  "from pinned host memory by the compaction kernel, D2H overlapped"}[mode],
        }

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if not args.no_cpu and world == 1:
        v, cores, sample, _ = cpu_leg(w, budget_s=20.0)
        cpu_baseline = {"value": v, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample}

    line = {
        "metric": metric, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config, "roofline": roofline,
        "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches * args.steps,
        "clocks": clocks.summary(), "token_heads_per_s": value * w["Hkv"],
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
