set -x
O=gpurun_out/r2o; mkdir -p $O
timeout 600 ncu --metrics launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__waves_per_multiprocessor,gpu__time_duration.sum,launch__shared_mem_per_block_dynamic,launch__shared_mem_per_block_static,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:rollout_fwd -s 1 -c 1 python tools/profile_target.py hand20-adjoint-bdf1-b2048 > $O/occ.log 2>&1; tail -15 $O/occ.log
