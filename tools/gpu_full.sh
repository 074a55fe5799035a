#!/bin/bash
# the round-end sequence on one GPU: every GPU test, smoke(), the bench (both arms); outputs under gpurun_out/<tag>_*
tag=${1:-r4}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | cut -c1-400 | tee gpurun_out/${tag}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/${tag}_smoke.txt
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc $?"; tail -c 1500 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "ref rc $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench.json").readline())
print({k: d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["roofline_step"]["frac"], d["e2e"], d.get("latency"), d.get("config4"), d["cpu_baseline"], d["check"])
print(open("gpurun_out/${tag}_bench_reference.json").readline()[:600])
PY
