set -x
O=gpurun_out/r2e; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --workload chain64-bdf1-b8192 --steps 3 --warmup 3 --no-cpu > $O/bench_chain64.log 2>&1
timeout 400 python bench.py --workload chain32-ground-bdf2-b4096 --steps 3 --warmup 3 --no-cpu > $O/bench_ground.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_default.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_fwd -s 1 -c 1 -o $O/ncu_ground python tools/profile_target.py chain32-ground-bdf2-b4096 > $O/ncu_ground.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_fwd -s 1 -c 1 -o $O/ncu_headline python tools/profile_target.py chain32-bdf1-b4096 > $O/ncu_headline.log 2>&1
ls -la $O
