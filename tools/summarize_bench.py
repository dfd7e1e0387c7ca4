#!/usr/bin/env python
"""Markdown rows from bench.py JSON lines: python tools/summarize_bench.py file.json [...]"""
import json
import sys


def last_json(path):
    for line in reversed(open(path).read().strip().splitlines()):
        line = line.strip()
        if line.startswith("{"):
            return json.loads(line)
    return None


for path in sys.argv[1:]:
    d = last_json(path)
    if d is None:
        print(f"| {path} | no JSON line |")
        continue
    c, r, e = d["config"], d.get("roofline") or {}, d.get("e2e") or {}
    eager = d.get("eager") or {}
    print(f"| {c['workload']} | {d['ms_per_step'] * 1e3:.1f} | {d['value'] / 1e9:.3f} G | "
          f"{r.get('algorithmic_bytes_per_launch', 0) / 1e6:.1f} | {r.get('achieved', 0):.0f} | {r.get('frac', 0):.3f} | "
          f"{d.get('host_us_per_call', float('nan')):.1f} | {eager.get('ms_per_step', float('nan')) * 1e3:.1f} / "
          f"{eager.get('host_us_per_call', float('nan')):.1f} | {e.get('ms_per_step', float('nan')):.2f} ({e.get('mode')}) | "
          f"{d['clocks'].get('sm_mhz')} |")
    for x in d.get("extras") or []:
        keys = [k for k in ("us_per_step", "ms_per_prefill_pass", "us_per_layer", "ms_per_call", "prologue_ms", "scan_ms",
                            "frac_of_hbm_peak", "frac_of_hbm_peak_rank0", "tokens_per_s", "host_us_per_call",
                            "eager_us_per_step", "eager_host_us_per_call", "hook_us_per_layer_token_no_compaction",
                            "compactions", "compaction_host_us_per_call", "compaction_device_us_per_call_mean",
                            "compaction_device_us_per_call_min", "generated_tokens") if k in x]
        print("  - extra", x["workload"], {k: (round(x[k], 3) if isinstance(x[k], float) else x[k]) for k in keys})
    if d.get("extras_errors"):
        print("  - extras_errors", d["extras_errors"])
    cb = d.get("cpu_baseline")
    if cb:
        print(f"  - cpu_baseline {cb['value']:.0f} {cb['unit']} on {cb['cores']} threads, kind {cb['kind']}: {cb['sample']}")
