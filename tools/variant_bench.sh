#!/bin/bash
# Run bench.py once per tuning variant in build/variants/ and print the kernel times (GPU box).
cd "$(dirname "$0")/.."
for so in build/variants/lib_*.so; do
  name=$(basename $so .so)
  SMH_B200_LIB=$PWD/$so timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 5 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['kernels']
print('$name', 'fps %.0f' % d['value'], 'px %.3f ms' % k['k_pixels']['ms'], 'lsd %.3f ms' % k['k_lsd']['ms'], 'frac %.3f' % d['roofline']['frac'], 'steps/frame %.0f' % k['k_lsd']['ray_steps_per_frame'])"
done
