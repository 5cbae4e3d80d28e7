set -x
O=gpurun_out/r2u; mkdir -p $O
timeout 900 python -m pytest tests/test_golden.py tests/test_gpu_drivers.py tests/test_gpu_euler_joints.py tests/test_gpu_parity.py -q -m gpu -x -n 4 > $O/tests.log 2>&1; tail -3 $O/tests.log
timeout 600 python tools/straggler_scaling.py > $O/straggler_scaling.log 2>&1; cat $O/straggler_scaling.log
