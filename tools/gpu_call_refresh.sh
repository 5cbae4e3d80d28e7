# Refresh of the numbers a round ends with: smoke, the GPU suite, every bench line (with the CPU arm), the launch list.
#   gpurun --timeout 2400 -- 'bash tools/gpu_call_refresh.sh'   ->  gpurun_out/refresh/*  (copied to profiles/r02_* by hand)
set -x
O=gpurun_out/refresh; mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 1200 python -m pytest tests -q -m gpu -n 4 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/bench_chain32.log 2>&1
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_chain32_reference.log 2>&1
timeout 400 python bench.py --workload chain10-bdf1-b1024 --steps 10 --warmup 3 > $O/bench_chain10.log 2>&1
timeout 600 python bench.py --workload chain32-ground-bdf2-b4096 --steps 5 --warmup 3 > $O/bench_ground.log 2>&1
timeout 600 python bench.py --workload hand20-adjoint-bdf1-b2048 --steps 10 --warmup 3 > $O/bench_adjoint.log 2>&1
timeout 900 python bench.py --workload chain64-bdf1-b8192 --steps 5 --warmup 3 > $O/bench_chain64.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_chain32.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_chain32_under_ncu.log 2>&1
ls -la $O
