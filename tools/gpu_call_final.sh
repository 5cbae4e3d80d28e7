# Final measurement pass of round 2: every bench line, the launch list and the ncu captures committed under profiles/.
set -x
O=gpurun_out/r2end; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt; nproc >> $O/smi.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/bench_chain32.log 2>&1
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_chain32_reference.log 2>&1
timeout 400 python bench.py --workload chain10-bdf1-b1024 --steps 10 --warmup 3 > $O/bench_chain10.log 2>&1
timeout 600 python bench.py --workload chain32-ground-bdf2-b4096 --steps 5 --warmup 3 > $O/bench_ground.log 2>&1
timeout 600 python bench.py --workload hand20-adjoint-bdf1-b2048 --steps 10 --warmup 3 > $O/bench_adjoint.log 2>&1
timeout 900 python bench.py --workload chain64-bdf1-b8192 --steps 5 --warmup 3 > $O/bench_chain64.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_chain32.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_chain32_under_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_adjoint.csv python bench.py --workload hand20-adjoint-bdf1-b2048 --steps 2 --warmup 3 --no-cpu > $O/bench_adjoint_under_ncu.log 2>&1
for w in chain32-bdf1-b4096 chain64-bdf1-b8192 chain32-ground-bdf2-b4096 chain10-bdf1-b1024; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_fwd -s 1 -c 1 -o $O/ncu_$w python tools/profile_target.py $w > $O/ncu_$w.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_fwd -s 1 -c 1 -o $O/ncu_adjoint_fwd python tools/profile_target.py hand20-adjoint-bdf1-b2048 > $O/ncu_adjoint_fwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:adjoint_bwd -s 1 -c 1 -o $O/ncu_adjoint_bwd python tools/profile_target.py hand20-adjoint-bdf1-b2048 > $O/ncu_adjoint_bwd.log 2>&1
timeout 900 python tools/explore_r2b.py > $O/stalled_newton_shortcuts.log 2>&1
timeout 300 python tools/ground_nocontact.py > $O/ground_nocontact.log 2>&1
ls -la $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_adjoint.py tests/test_gpu_long_chains.py tests/test_mex_gateway.py tests/test_gpu_trees.py -m gpu -q -x -k "scene100_101 or hand_c4 or newton_system_two_warps or eval_long_chain or gateway or lockstep or tree_adjoint" > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_long_chains.py tests/test_gpu_adjoint.py -m gpu -q -x -k "newton_system_two_warps or lockstep or hand_c4" > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck.log
ls -la $O
