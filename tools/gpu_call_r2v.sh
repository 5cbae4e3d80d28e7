set -x
O=gpurun_out/r2v; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -n 4 > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_long_chains.py tests/test_gpu_euler_joints.py -m gpu -q -x -k "schedule or lockstep or scene7" > $O/memcheck_sched.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck_sched.log; tail -5 $O/memcheck_sched.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench.log 2>&1; tail -1 $O/bench.log | cut -c1-220
