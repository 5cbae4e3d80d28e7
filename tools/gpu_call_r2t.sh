set -x
O=gpurun_out/r2t; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_long_chains.py -q -m gpu -x -k lockstep > $O/tests_lockstep.log 2>&1; tail -12 $O/tests_lockstep.log
