#!/bin/bash
# Developer helper (GPU box): for every tagged library variant run the LU-sensitive parity tests and the quick timing probe.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for so in redmax_b200/lib/libredmax_b200_*.so; do
  tag=$(basename $so .so); tag=${tag#libredmax_b200_}
  echo "=== variant $tag" 
  RMX_LIB=$PWD/$so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
  RMX_LIB=$PWD/$so timeout 200 python tools/quick_bench.py 2>&1 | grep -v "^NVIDIA"
done
