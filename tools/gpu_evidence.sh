#!/bin/bash
# Final evidence of a round on one GPU.  usage: tools/gpu_evidence.sh <tag>
#   <tag>_launches.csv/.txt   ncu launch list of two bench steps (duration, DRAM bytes, warp instructions per launch)
#   <tag>_<kernel>.ncu-rep + _details.txt + _raw.csv + _sass.csv   ncu --set full of each hot launch (1024-frame batch)
#   <tag>_sanitize_*.txt      compute-sanitizer sweep (tools/sanitize.sh)
tag=${1:-r4}
mkdir -p gpurun_out
bash tools/gpu_launches.sh ${tag} > gpurun_out/${tag}_launches.txt 2>&1; cat gpurun_out/${tag}_launches.txt | cut -c1-260
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_pixels2|^k_px_seams|^k_lsd2" -c 3 -s 6 -f -o gpurun_out/${tag}_hot \
   python bench.py --frames 1024 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --serial > gpurun_out/${tag}_hot_ncu.log 2>&1
ncu -i gpurun_out/${tag}_hot.ncu-rep --page details > gpurun_out/${tag}_hot_details.txt 2>&1
ncu -i gpurun_out/${tag}_hot.ncu-rep --page raw --csv > gpurun_out/${tag}_hot_raw.csv 2>&1
ncu -i gpurun_out/${tag}_hot.ncu-rep --page source --csv > gpurun_out/${tag}_hot_sass.csv 2>&1
grep -E "^  [a-z].*\(|Duration|Executed Ipc Active|Issue Slots Busy|DRAM Throughput|Executed Instructions  |Achieved Occupancy|Registers Per" gpurun_out/${tag}_hot_details.txt
[ "$2" = "nosan" ] || bash tools/sanitize.sh ${tag}
