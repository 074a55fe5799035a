#!/usr/bin/env python3
"""Per-source-line hot spots from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
usage: ncu_lines.py dump.csv [top]   (first launch in the dump only)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []
hdr = None
seen_files = 0
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Function Name":
        seen_files += 1
        continue
    if hdr is None or not r or not r[0].isdigit():
        continue
    col = {n: i for i, n in enumerate(hdr)}
    W = col["Warp Stall Sampling (All Samples)"]
    I = col["Instructions Executed"]
    try:
        stalls = {n[6:]: int(r[i]) for n, i in col.items() if n.startswith("stall_") and "Not Issued" not in n and r[i].isdigit()}
        out.append((cur_file, int(r[0]), r[1].strip(), int(r[W]), int(r[I]), stalls))
    except (ValueError, IndexError):
        pass
# the dump repeats per launch: keep the first occurrence of each (file, line)
first = {}
for o in out:
    first.setdefault((o[0], o[1]), o)
out = list(first.values())
tw = sum(o[3] for o in out) or 1
ti = sum(o[4] for o in out) or 1
print(f"samples {tw} instructions {ti}")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    why = ", ".join(f"{k} {v}" for k, v in sorted(o[5].items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{o[3] / tw * 100:5.1f}% stall {o[4] / ti * 100:5.1f}% inst  {o[0]}:{o[1]:<4d} {o[2][:100]}  [{why}]")
