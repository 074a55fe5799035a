#!/bin/bash
# ncu --set full of one kernel on the bench launch + source dumps.  usage: tools/gpu_ncu_one.sh <tag> <kernel regex> [bench args]
tag=$1; k=$2; shift; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$k" -c 1 -s 3 -f -o gpurun_out/${tag} \
   python bench.py --frames 1024 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --serial "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page details > gpurun_out/${tag}_details.txt 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_sass.csv 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${tag}_src.csv 2>&1
grep -E "Duration|Executed Ipc Active|Issue Slots Busy|Registers Per|Achieved Occupancy|Executed Instructions|Warp Cycles Per Issued" gpurun_out/${tag}_details.txt
