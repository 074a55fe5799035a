#!/bin/bash
# bench the tuning variants (build/variants/*.so) and the CTAs-per-SM cap of the line search
run() { timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['kernels']
print('$1', 'fps %.0f' % d['value'], 'step %.3f' % d['ms_per_step'], 'px %.3f ms' % k['k_pixels2']['ms'], 'lsd %.3f ms' % k['k_lsd2']['ms'])"; }
run base
for c in 3 4 5 6; do SMH_B200_LSD_CTAS_PER_SM=$c run ctas$c; done
for so in build/variants/lib_nw2.so build/variants/lib_nw8.so build/variants/lib_band128.so build/variants/lib_band32.so; do
  SMH_B200_LIB=$PWD/$so run $(basename $so .so)
done
SMH_B200_LIB=$PWD/build/variants/lib_nw2.so SMH_B200_LSD_CTAS_PER_SM=8 run nw2_ctas8
SMH_B200_LIB=$PWD/build/variants/lib_nw2.so SMH_B200_LSD_CTAS_PER_SM=6 run nw2_ctas6
