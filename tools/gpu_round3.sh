#!/bin/bash
# Full GPU-box visit: parity tests, smoke, bench (both arms), launch lists, ncu --set full of the hot kernels,
# the reference's own CUDA kernels timed on the same box.  usage: tools/gpu_round3.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 1200 gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench_reference.json
# launch list of the bench command (kernels serialised by ncu: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
grep -c . gpurun_out/${tag}_launches.csv
for k in k_pixels k_lsd; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k" -c 1 -s 3 -f -o gpurun_out/${tag}_${k} \
     python bench.py --frames 256 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --serial > gpurun_out/${tag}_ncu_${k}.log 2>&1
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page details > gpurun_out/${tag}_${k}_details.txt 2>&1
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>&1
done
timeout 300 python tools/refgpu_time.py 16 > gpurun_out/${tag}_refgpu_time.json 2>&1; tail -2 gpurun_out/${tag}_refgpu_time.json
ls gpurun_out | grep ${tag} | wc -l
