set -x
O=gpurun_out/r2z2; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -k "gpus or shard or ngpu or two_gpu or multi" > $O/pytest_2gpu.log 2>&1; tail -4 $O/pytest_2gpu.log
