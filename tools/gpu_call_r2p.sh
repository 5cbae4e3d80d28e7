set -x
O=gpurun_out/r2p; mkdir -p $O
w=chain32-ground-bdf2-b4096
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_fwd -s 1 -c 1 -o $O/ncu_$w python tools/profile_target.py $w > $O/ncu_$w.log 2>&1
ls -la $O
