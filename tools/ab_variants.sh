#!/bin/bash
# Developer helper (GPU box): timing probe of the default library and of every tagged variant (tools/build_variant.sh);
# with RMX_AB_TESTS=1 also the forward parity tests per variant.
cd "$(dirname "$0")/.."
echo "=== default"; timeout 200 python tools/quick_bench.py 2>&1 | grep -v "^NVIDIA"
for so in redmax_b200/lib/libredmax_b200_*.so; do
  [ -e "$so" ] || continue
  tag=$(basename $so .so); tag=${tag#libredmax_b200_}
  echo "=== variant $tag"
  if [ -n "$RMX_AB_TESTS" ]; then RMX_LIB=$PWD/$so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2; fi
  RMX_LIB=$PWD/$so timeout 200 python tools/quick_bench.py 2>&1 | grep -v "^NVIDIA"
done
