# usage: bash tools/gpu_call_multi.sh N   (inside gpurun --gpus N)
set -x
N=$1
O=gpurun_out/r2m; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/smi_$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if [ "$N" = "8" ]; then
  timeout 900 $TR bench.py --gpus $N --workload chain64-bdf1-b8192 --steps 3 --warmup 3 > $O/bench_chain64_weak_${N}gpu.log 2>&1
fi
if [ "$N" = "4" ]; then
  timeout 600 $TR bench.py --gpus $N --workload hand20-adjoint-bdf1-b2048 --steps 5 --warmup 3 > $O/bench_adjoint_weak_${N}gpu.log 2>&1
fi
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_headline_weak_${N}gpu.log 2>&1
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --scaling strong > $O/bench_headline_strong_${N}gpu.log 2>&1
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests -m gpu -q -k "gpus or shard or ngpu or two_gpu or multi" > $O/pytest_2gpu.log 2>&1
fi
ls -la $O
