#!/usr/bin/env python3
"""Summarise an `ncu -i X.ncu-rep --page source --csv` dump: hottest SASS ranges by stall
samples / executed instructions, and the stall-reason mix.  usage: ncu_hot.py dump.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
S, I, T, W = col["Source"], col["Instructions Executed"], col["Thread Instructions Executed"], col["Warp Stall Sampling (All Samples)"]
stall_cols = [(n, i) for n, i in col.items() if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[I]), int(r[W]), int(r[T]), r[S].strip(), {n: int(r[i]) for n, i in stall_cols}))
    except (ValueError, IndexError):
        pass
ti = sum(d[0] for d in data) or 1
tw = sum(d[1] for d in data) or 1
print(f"instructions {ti}  samples {tw}  avg threads/inst {sum(d[2] for d in data) / ti:.1f}")
mix = {}
for d in data:
    for n, v in d[4].items():
        mix[n] = mix.get(n, 0) + v
print("stall mix:", ", ".join(f"{n[6:]} {v / tw * 100:.1f}%" for n, v in sorted(mix.items(), key=lambda kv: -kv[1])[:8]))
print(f"{'idx':>5} {'inst%':>6} {'stall%':>6} {'thr':>4}  sass")
order = sorted(range(len(data)), key=lambda k: -data[k][1])[:top]
for k in sorted(order):
    d = data[k]
    reason = max(d[4].items(), key=lambda kv: kv[1])[0][6:] if d[1] else ""
    print(f"{k:5d} {d[0] / ti * 100:6.2f} {d[1] / tw * 100:6.2f} {d[2] / max(d[0], 1):4.0f}  {d[3][:90]}  [{reason}]")
