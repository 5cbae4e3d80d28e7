set -x
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_default.log 2>&1
RMX_LIB=$PWD/build/lib_v200/libredmax_b200.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_default_v200.log 2>&1
timeout 300 python bench.py --workload chain10-bdf1-b1024 --steps 5 --warmup 3 --no-cpu > $O/bench_chain10.log 2>&1
RMX_LIB=$PWD/build/lib_v200/libredmax_b200.so timeout 300 python bench.py --workload chain10-bdf1-b1024 --steps 5 --warmup 3 --no-cpu > $O/bench_chain10_v200.log 2>&1
timeout 600 python bench.py --workload chain64-bdf1-b8192 --steps 3 --warmup 3 --no-cpu > $O/bench_chain64.log 2>&1
timeout 300 python bench.py --workload hand20-adjoint-bdf1-b2048 --steps 5 --warmup 3 --no-cpu > $O/bench_adjoint.log 2>&1
timeout 400 python bench.py --workload chain32-ground-bdf2-b4096 --steps 3 --warmup 3 --no-cpu > $O/bench_ground.log 2>&1
RMX_LIB=$PWD/build/lib_v200/libredmax_b200.so timeout 400 python bench.py --workload chain32-ground-bdf2-b4096 --steps 3 --warmup 3 --no-cpu > $O/bench_ground_v200.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_fwd -s 1 -c 1 -o $O/ncu_chain64 python tools/profile_target.py chain64-bdf1-b8192 > $O/ncu_chain64.log 2>&1
ls -la $O
