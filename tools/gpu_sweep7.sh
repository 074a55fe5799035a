#!/bin/bash
# overlap of the pixel pass and the line search: batches in flight x persistent line-search CTAs per SM
run() { timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 6 $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['kernels']
print('$1', 'fps %.0f' % d['value'], 'step %.3f' % d['ms_per_step'], 'px %.3f ms' % k['k_pixels2']['ms'], 'lsd %.3f ms' % k['k_lsd2']['ms'])"; }
for f in 1 2 3 4; do run inflight$f "--in-flight $f"; done
for c in 3 4 5 6; do SMH_B200_LSD_CTAS_PER_SM=$c run ctas${c}_if2 "--in-flight 2"; SMH_B200_LSD_CTAS_PER_SM=$c run ctas${c}_if3 "--in-flight 3"; done
