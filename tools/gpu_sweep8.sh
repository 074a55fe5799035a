#!/bin/bash
run() { timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['kernels']
print('$1', 'fps %.0f' % d['value'], 'step %.3f' % d['ms_per_step'], 'px %.3f ms' % k['k_pixels2']['ms'], 'lsd %.3f ms' % k['k_lsd2']['ms'])"; }
run base
for so in build/variants/lib_*.so; do
  for smem in -1 0; do for c in -2 50 70; do
    SMH_B200_PX_SMEM=$smem SMH_B200_PX_CARVEOUT=$c SMH_B200_LIB=$PWD/$so run "$(basename $so .so) smem=$smem carve=$c"
  done; done
done
