#!/bin/bash
# launch list of one bench step (ncu, gpu__time_duration + dram bytes per launch; cold-cache, serialised)
tag=${1:-r4}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -s 6 -c 16 --csv \
   --log-file gpurun_out/${tag}_launches.csv python bench.py --frames 1024 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --serial > gpurun_out/${tag}_launches.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/${tag}_launches.csv")))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
hdr=rows[hi]; c={n:i for i,n in enumerate(hdr)}
cur={}
for r in rows[hi+1:]:
    key=(r[c["ID"]], r[c["Kernel Name"]][:60], r[c["Grid Size"]])
    cur.setdefault(key,{})[r[c["Metric Name"]]]=(r[c["Metric Value"]], r[c["Metric Unit"]])
for k,v in cur.items():
    print(k[0], k[1], k[2], " ".join(f"{m.split('.')[0].replace('__','_')}={x[0]}{x[1]}" for m,x in v.items()))
PY
