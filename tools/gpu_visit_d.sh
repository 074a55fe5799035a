#!/bin/bash
# tests given + full bench (both arms).  usage: tools/gpu_visit_d.sh <tag> [pytest args]
tag=${1:-r3x}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest "$@" -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -25 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench.json").readline()); k=d["kernels"]
    print("fps %.0f step %.3f ms px %.3f lsd %.3f frac %.3f step_frac %.3f" % (d["value"], d["ms_per_step"], k["k_pixels2"]["ms"], k["k_lsd2"]["ms"], d["roofline"]["frac"], d["roofline_step"]["frac"]))
    e=d["e2e"]; print("e2e %.0f (l8 %.0f) ceiling %.0f frac %.2f match %s" % (e["value"], e["ui_l8"]["value"], e["bus"]["ceiling_frames_per_s"], e["bus"]["frac_of_ceiling"], e["lines_match_oracle"]))
    print("latency", json.dumps(d["latency"])[:900])
    print("config4", json.dumps(d["config4"])[:600])
    print("check", d["check"], "cpu", d["cpu_baseline"])
except Exception as ex:
    print("bench parse failed", ex)
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench_reference.json | cut -c1-400
