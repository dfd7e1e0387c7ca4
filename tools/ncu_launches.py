#!/usr/bin/env python
"""Summarise an `ncu --csv` launch list (gpu__time_duration + dram bytes) per launch and per kernel."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii, gi = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size"))
d = defaultdict(dict)
for r in rows[1:]:
    d[(int(r[ii]), r[ki].split("(")[0][-48:], r[gi])][r[mi]] = float(r[vi].replace(",", ""))
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for (i, k, g), m in sorted(d.items()):
    t = m.get("gpu__time_duration.sum", 0) / 1e3
    rd, wr = m.get("dram__bytes_read.sum", 0) / 1e6, m.get("dram__bytes_write.sum", 0) / 1e6
    print(f"{i:4d} {k:48s} grid {g:14s} {t:8.1f} us  read {rd:8.1f} MB  write {wr:8.1f} MB  {(rd + wr) / max(t, 1e-9) / 1e3:6.2f} TB/s")
    a = agg[k]
    a[0] += 1; a[1] += t; a[2] += rd; a[3] += wr
tot = sum(a[1] for a in agg.values())
print("---- per kernel (avg per launch, share of listed time)")
for k, a in agg.items():
    print(f"{k:48s} n={a[0]:3d} {a[1] / a[0]:8.1f} us  read {a[2] / a[0]:8.1f} MB write {a[3] / a[0]:8.1f} MB  share {100 * a[1] / tot:5.1f}%")

# optional: python tools/ncu_launches.py launches.csv <workload> -> update profiles/traffic.json with the
# per-call DRAM traffic (kernels of one compress call = one launch of each listed kvp kernel)
if len(sys.argv) > 2:
    import json, os
    per_call = {k: (a[2] + a[3]) / a[0] * 1e6 for k, a in agg.items() if "at::" not in k and "elementwise" not in k}
    times = {k: round(a[1] / a[0], 1) for k, a in agg.items() if "at::" not in k and "elementwise" not in k}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[sys.argv[2]] = {"dram_bytes_per_call": int(sum(per_call.values())), "kernel_us": times,
                         "source": "ncu launch list " + os.path.basename(sys.argv[1]) + " (dram__bytes_read.sum + dram__bytes_write.sum, --cache-control none)"}
    json.dump(data, open(path, "w"), indent=1)
    print("traffic.json updated for", sys.argv[2], data[sys.argv[2]]["dram_bytes_per_call"] / 1e6, "MB")
