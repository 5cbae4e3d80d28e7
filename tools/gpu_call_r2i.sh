set -x
O=gpurun_out/r2i; mkdir -p $O
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_default.log 2>&1
timeout 600 python bench.py --workload chain64-bdf1-b8192 --steps 3 --warmup 3 --no-cpu > $O/bench_chain64.log 2>&1
timeout 300 python bench.py --workload hand20-adjoint-bdf1-b2048 --steps 5 --warmup 3 --no-cpu > $O/bench_adjoint.log 2>&1
timeout 300 python bench.py --workload chain10-bdf1-b1024 --steps 5 --warmup 3 --no-cpu > $O/bench_chain10.log 2>&1
timeout 400 python bench.py --workload chain32-ground-bdf2-b4096 --steps 3 --warmup 3 --no-cpu > $O/bench_ground.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
ls -la $O
