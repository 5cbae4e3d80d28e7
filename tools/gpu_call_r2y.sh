set -x
O=gpurun_out/r2y; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_edges.py -q -m gpu -n 4 > $O/edges.log 2>&1; tail -25 $O/edges.log
