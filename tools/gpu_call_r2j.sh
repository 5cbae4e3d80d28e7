set -x
O=gpurun_out/r2j; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_default.log 2>&1
timeout 600 python bench.py --workload chain64-bdf1-b8192 --steps 3 --warmup 3 --no-cpu > $O/bench_chain64.log 2>&1
timeout 300 python bench.py --workload hand20-adjoint-bdf1-b2048 --steps 5 --warmup 3 --no-cpu > $O/bench_adjoint.log 2>&1
timeout 300 python bench.py --workload chain10-bdf1-b1024 --steps 5 --warmup 3 --no-cpu > $O/bench_chain10.log 2>&1
timeout 600 python tools/bench_linsolve.py > $O/linsolve_lu_vs_pcg.log 2>&1
ls -la $O
