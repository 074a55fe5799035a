#!/bin/bash
# comparator + sanitizer visit.  usage: tools/gpu_visit_e.sh <tag>
tag=${1:-r3e}
mkdir -p gpurun_out
timeout 600 python tools/refgpu_time.py 256 > gpurun_out/${tag}_refgpu_time.json 2> gpurun_out/${tag}_refgpu_time.err; tail -c 1800 gpurun_out/${tag}_refgpu_time.json
# DRAM bytes + duration of every reference launch of one frame
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
   --log-file gpurun_out/${tag}_refgpu_ncu.csv python tools/refgpu_time.py --one > gpurun_out/${tag}_refgpu_ncu.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/${tag}_refgpu_ncu.csv")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
col = {n: i for i, n in enumerate(rows[hi])}
acc = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= col["Metric Value"]: continue
    k = r[col["Kernel Name"]].split("(")[0]
    a = acc.setdefault(k, {"n": 0})
    a[r[col["Metric Name"]]] = a.get(r[col["Metric Name"]], 0.0) + float(r[col["Metric Value"]].replace(",", ""))
    if r[col["Metric Name"]] == "gpu__time_duration.sum": a["n"] += 1
with open("gpurun_out/${tag}_refgpu_ncu_summary.txt", "w") as f:
    f.write("reference kernels, ONE 2560x1440 frame (ncu: per-launch metrics summed per kernel)\n")
    f.write("%-55s %8s %12s %14s %14s\n" % ("kernel", "launches", "time us", "dram read B", "dram write B"))
    for k, a in acc.items():
        f.write("%-55s %8d %12.1f %14.0f %14.0f\n" % (k[:55], a["n"], a.get("gpu__time_duration.sum", 0) / 1e3, a.get("dram__bytes_read.sum", 0), a.get("dram__bytes_write.sum", 0)))
print(open("gpurun_out/${tag}_refgpu_ncu_summary.txt").read())
PY
bash tools/sanitize.sh ${tag}
