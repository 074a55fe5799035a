#!/bin/bash
# ncu --set full of the first launches matching a regex (both pixel-pass launches), details + source dumps per launch
tag=$1; k=$2; cnt=${3:-2}; shift; shift; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$k" -c $cnt -s 6 -f -o gpurun_out/${tag} \
   python bench.py --frames 1024 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --serial "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page details > gpurun_out/${tag}_details.txt 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_sass.csv 2>&1
grep -E "^  [a-z].*\(|Duration|Executed Ipc Active|Issue Slots Busy|DRAM Throughput|Executed Instructions|Warp Cycles Per Issued|Achieved Occupancy|L2 Hit" gpurun_out/${tag}_details.txt
