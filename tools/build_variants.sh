#!/bin/bash
# Build tuning variants of libsmh_b200.so into build/variants/ (travels to the GPU box; git-ignored).
# usage: tools/build_variants.sh name1:"-DFOO=1 -DBAR=2" name2:"..."
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  (cd squad-mortar-helper_b200/csrc && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -fmad=false -Xcompiler -fPIC,-ffp-contract=off -shared -cudart static -ccbin /usr/bin/g++ $flags \
     -Xptxas=-v -o ../../build/variants/lib_$name.so smh_b200.cu smh_host_widen.cpp 2>&1 | grep -A2 "k_lsdILb0\|k_pixelsILb1" | grep "Used\|spill" | sed "s/^/[$name] /") &
done
wait
ls -la build/variants
