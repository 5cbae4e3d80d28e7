set -x
O=gpurun_out/r2l; mkdir -p $O
timeout 400 python bench.py --workload chain32-ground-bdf2-b4096 --steps 3 --warmup 3 --no-cpu > $O/bench_ground.log 2>&1
RMX_LIB=$PWD/build/lib_gunroll/libredmax_b200.so timeout 400 python bench.py --workload chain32-ground-bdf2-b4096 --steps 3 --warmup 3 --no-cpu > $O/bench_ground_unroll.log 2>&1
ls -la $O
