set -x
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
nproc >> $O/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_default.log 2>&1
RMX_LIB=$PWD/build/lib_rcp/libredmax_b200.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_rcp_early.log 2>&1
timeout 300 python bench.py --workload hand20-adjoint-bdf1-b2048 --steps 5 --warmup 3 > $O/bench_adjoint.log 2>&1
timeout 300 python bench.py --workload chain10-bdf1-b1024 --steps 5 --warmup 3 --no-cpu > $O/bench_chain10.log 2>&1
timeout 400 python bench.py --workload chain32-ground-bdf2-b4096 --steps 3 --warmup 3 --no-cpu > $O/bench_ground.log 2>&1
timeout 600 python bench.py --workload chain64-bdf1-b8192 --steps 2 --warmup 3 --no-cpu > $O/bench_chain64.log 2>&1
timeout 600 python tools/explore_r2.py > $O/explore.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:adjoint_bwd -s 1 -c 1 -o $O/ncu_adjoint_bwd python tools/profile_target.py hand20-adjoint-bdf1-b2048 > $O/ncu_adjoint_bwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_fwd -s 1 -c 1 -o $O/ncu_adjoint_fwd python tools/profile_target.py hand20-adjoint-bdf1-b2048 > $O/ncu_adjoint_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_fwd -s 1 -c 1 -o $O/ncu_ground python tools/profile_target.py chain32-ground-bdf2-b4096 > $O/ncu_ground.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_adjoint.csv python bench.py --workload hand20-adjoint-bdf1-b2048 --steps 2 --warmup 3 --no-cpu > $O/bench_adjoint_under_ncu.log 2>&1
ls -la $O
