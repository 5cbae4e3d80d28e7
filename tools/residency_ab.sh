# Developer experiment: throughput of the forward kernels against the number of co-resident blocks per SM, set by padding the
# dynamic shared memory of each block (RMX_DEBUG_SMEM_PAD).  Tells whether a layout that fits one more block per SM would pay.
set -x
O=gpurun_out/residency; mkdir -p $O
run() { # workload pad tag
  RMX_DEBUG_SMEM_PAD=$2 timeout 300 python bench.py --workload $1 --steps 6 --warmup 3 --no-cpu > $O/$1.$3.log 2>&1
  python - $O/$1.$3.log $1 $3 $2 <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('RESIDENCY %-28s blocks/SM %s pad %6s B  %8.3f ms  %.3f M rollout-steps/s' % (sys.argv[2], sys.argv[3], sys.argv[4], d['ms_per_step'], d['value']/1e6))
PY
}
run chain32-ground-bdf2-b4096 0 5
run chain32-ground-bdf2-b4096 6144 4
run chain32-ground-bdf2-b4096 17408 3
run chain32-bdf1-b4096 0 8
run chain32-bdf1-b4096 12288 6
run chain32-bdf1-b4096 25600 4
run hand20-adjoint-bdf1-b2048 0 8
run hand20-adjoint-bdf1-b2048 5120 6
