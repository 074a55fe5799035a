#!/bin/bash
# quick parity + timing visit.  usage: tools/gpu_visit_c.sh <tag> [pytest args]
tag=${1:-r3c}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest "$@" -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -25 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/${tag}_bench_quick.json 2> gpurun_out/${tag}_bench_quick.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_quick.json").readline()); k=d["kernels"]
    print("fps %.0f  step %.3f ms  px %.3f ms  lsd %.3f ms  frac %.3f  seeds/frame %.1f steps/frame %.0f" % (d["value"], d["ms_per_step"], k["k_pixels2"]["ms"], k["k_lsd2"]["ms"], d["roofline"]["frac"], k["k_lsd2"]["seeds_per_frame"], k["k_lsd2"]["ray_steps_per_frame"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_bench_quick.err").read()[-2000:])
PY
