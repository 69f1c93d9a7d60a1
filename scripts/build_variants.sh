#!/bin/bash
# builds of libfn3 with different compile-time knobs, for A/B timing on the GPU box: scripts/build_variants.sh name "-DX=1 ..." ...
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -I include $flags \
       -o variants/libfn3_$name.so findneighbour3_b200/csrc/fn3_engine.cu &
done
wait
ls -la variants/
