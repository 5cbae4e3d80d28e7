set -x
O=gpurun_out/final_check; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -n 4 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
for w in chain32-bdf1-b4096 chain32-ground-bdf2-b4096; do
  timeout 300 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu > $O/bench_$w.log 2>&1
  python - $O/bench_$w.log $w <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('BENCH %-28s %8.3f ms  %.3f M rollout-steps/s e2e %.3f pageable %.3f' % (sys.argv[2], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['e2e_pageable']['value']/1e6))
PY
done
