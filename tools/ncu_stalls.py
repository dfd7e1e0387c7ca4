#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump: stall reasons overall and the hottest SASS lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
body = rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
for r in body:
    for s in stalls:
        try:
            tot[s] += int(r[idx[s]])
        except (ValueError, IndexError):
            pass
all_s = sum(tot.values())
print("stall samples by reason:")
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:28s} {v:8d}  {100 * v / max(all_s, 1):5.1f}%")
print("hottest instructions (#samples):")
si = idx["# Samples"]
top = sorted(body, key=lambda r: -int(r[si] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    reasons = sorted(((int(r[idx[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"  {int(r[si]):7d}  {r[idx['Source']].strip()[:70]:70s} {reasons}")
