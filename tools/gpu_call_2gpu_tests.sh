set -x
O=gpurun_out/r2k; mkdir -p $O
NCCL_DEBUG=WARN timeout 600 python -m pytest tests -m gpu -q -k "gpus or shard or ngpu or two_gpu or multi" > $O/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_2gpu.log
tail -30 $O/pytest_2gpu.log
