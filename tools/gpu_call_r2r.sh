set -x
O=gpurun_out/r2r; mkdir -p $O
M=gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for g in 5 4; do
  RMX_GROUP=$g timeout 600 ncu --metrics $M --clock-control none -k regex:rollout_fwd -s 1 -c 1 python tools/profile_target.py chain32-ground-bdf2-b4096 > $O/icc_ground_g$g.log 2>&1; grep -A16 "Metric Name" $O/icc_ground_g$g.log
done
