set -x
O=gpurun_out/r2y; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -n 4 -x > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
for st in 1 0; do
for w in chain32-bdf1-b4096 chain64-bdf1-b8192; do
  RMX_STAGE=$st timeout 300 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu > $O/bench_${w}_stage$st.log 2>&1
  python - $O/bench_${w}_stage$st.log $w $st <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('STAGE=%s %-22s device %.3f ms  e2e pinned %.3f M  e2e pageable %.3f M' % (sys.argv[3], sys.argv[2], d['ms_per_step'], d['e2e']['value']/1e6, d['e2e_pageable']['value']/1e6))
PY
done; done
