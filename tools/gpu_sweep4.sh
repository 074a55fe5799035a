#!/bin/bash
# pixel pass: strips per CTA (SMH_B200_PX_GROUP) x rows per CTA (SMH_B200_PX_BAND)
run() { timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['kernels']
print('$1', 'fps %.0f' % d['value'], 'step %.3f' % d['ms_per_step'], 'px %.3f ms' % k['k_pixels2']['ms'], 'lsd %.3f ms' % k['k_lsd2']['ms'])"; }
run base
for g in 1 2 3 4 6 8 11; do SMH_B200_PX_GROUP=$g run group$g; done
for b in 32 128 256; do SMH_B200_PX_BAND=$b run band$b; done
