#!/bin/bash
# N-GPU visit: the multi-GPU tests, then the bench as the driver launches it.  usage: tools/gpu_multi.sh <tag> <N>
tag=$1; N=$2
mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest_multi_${N}gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
   > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err; echo "bench rc $?"; tail -c 600 gpurun_out/${tag}_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_${N}gpu.json").readline())
print({k: d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches")}, "frac", d["roofline"]["frac"], "\ne2e", {k: d["e2e"][k] for k in ("value","per_rank_frames_per_s")}, d["e2e"]["bus"]["ceiling_frames_per_s"], d["e2e"]["bus"]["frac_of_ceiling"], "\nper_rank", d["per_rank"]["ms_per_step"], "\nconfig4", d.get("config4"))
PY
