#!/bin/bash
# gpu_round.sh plus one `ncu --set full` capture of each hot kernel.  usage: tools/gpu_round_full.sh <tag>
tag=${1:-run}
bash tools/gpu_round.sh $tag
for k in k_pixels k_lsd; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k" -c 1 -s 3 -f -o gpurun_out/${tag}_${k} \
     python bench.py --frames 256 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_${k}.log 2>&1
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page details > gpurun_out/${tag}_${k}_details.txt 2>&1
done
ls -la gpurun_out | tail -20
