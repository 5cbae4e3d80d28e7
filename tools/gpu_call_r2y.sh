set -x
O=gpurun_out/r2y; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "load_balanced" > $O/sched_tests.log 2>&1; tail -12 $O/sched_tests.log
