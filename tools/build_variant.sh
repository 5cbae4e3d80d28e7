#!/bin/bash
# Developer helper: build a tagged variant of the CUDA library with extra -D flags (A/B experiments on the GPU box).
#   tools/build_variant.sh <tag> [-DFLAG ...]   ->  redmax_b200/lib/libredmax_b200_<tag>.so  (select with RMX_LIB=<path>)
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=true -Xcompiler -fPIC -shared -Xptxas -v \
  -I include "$@" -o redmax_b200/lib/libredmax_b200_${tag}.so redmax_b200/csrc/rmx_api.cu -lcudart > redmax_b200/lib/build_${tag}.log 2>&1
echo "built $tag"
