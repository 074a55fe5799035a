#!/bin/bash
# Evidence visit: per-phase cycles of k_lsd (LSD_PROFILE build), ncu --set full of both hot kernels on the
# 1024-frame bench launch with source counters, source/SASS dumps.  usage: tools/gpu_visit_a.sh <tag>
tag=${1:-r3a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/${tag}_bench_quick.json 2> gpurun_out/${tag}_bench_quick.err
tail -c 600 gpurun_out/${tag}_bench_quick.json
for k in k_lsd2 k_pixels2; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$k" -c 1 -s 3 -f -o gpurun_out/${tag}_${k} \
     python bench.py --frames 1024 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --serial > gpurun_out/${tag}_ncu_${k}.log 2>&1
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page details > gpurun_out/${tag}_${k}_details.txt 2>&1
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>&1
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page source --csv > gpurun_out/${tag}_${k}_sass.csv 2>&1
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${tag}_${k}_src.csv 2>&1
done
ls -la gpurun_out | grep ${tag}
