set -x
O=gpurun_out/r2n; mkdir -p $O
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_default.log 2>&1
RMX_LIB=$PWD/build/lib_v224/libredmax_b200.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_default_v224.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
ls -la $O
