set -x
O=gpurun_out/r2r; mkdir -p $O
M=gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for w in chain32-ground-bdf2-b4096 chain32-bdf1-b4096 hand20-adjoint-bdf1-b2048; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:rollout_fwd -s 1 -c 1 python tools/profile_target.py $w > $O/icc_$w.log 2>&1; grep -A14 "Metric Name" $O/icc_$w.log
done
