#!/bin/bash
run() { timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['kernels']
print('$1', 'fps %.0f' % d['value'], 'step %.3f' % d['ms_per_step'], 'px %.3f ms' % k['k_pixels2']['ms'], 'lsd %.3f ms' % k['k_lsd2']['ms'])"; }
run base
run base
SMH_B200_PX_NOSEAMS=1 run noseams
SMH_B200_PX_NOSEAMS=1 SMH_B200_LIB=$PWD/build/variants/lib_probe3.so run probe3_noseams
SMH_B200_LIB=$PWD/build/variants/lib_probe3.so run probe3
