#!/bin/bash
# pixel pass: one launch with the two kinds of CTAs interleaved (default) against two launches
run() { timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['kernels']
print('$1', 'fps %.0f' % d['value'], 'step %.3f' % d['ms_per_step'], 'px %.3f ms' % k['k_pixels2']['ms'], 'lsd %.3f ms' % k['k_lsd2']['ms'])"; }
run interleaved
SMH_B200_PX_MERGED=0 run two_launches
run interleaved
