set -x
O=gpurun_out/r2w; mkdir -p $O
timeout 600 python tools/oversubscribed_probe.py > $O/oversub.log 2>&1; cat $O/oversub.log
