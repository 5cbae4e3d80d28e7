set -x
O=gpurun_out/r2c; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_adjoint.py tests/test_gpu_long_chains.py tests/test_gpu_parity.py tests/test_mex_gateway.py -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --workload hand20-adjoint-bdf1-b2048 --steps 5 --warmup 3 --no-cpu > $O/bench_adjoint.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_default.log 2>&1
timeout 300 python tools/noise_probe.py > $O/noise.log 2>&1
RMX_IMPL=1 timeout 300 python tools/noise_probe.py >> $O/noise.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_fwd -s 1 -c 1 -o $O/ncu_adjoint_fwd python tools/profile_target.py hand20-adjoint-bdf1-b2048 > $O/ncu_adjoint_fwd.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu_all.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_all.log
ls -la $O
