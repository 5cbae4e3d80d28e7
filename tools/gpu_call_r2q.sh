set -x
O=gpurun_out/r2q; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -x -n 4 > $O/tests.log 2>&1; tail -4 $O/tests.log
for w in chain32-bdf1-b4096 chain32-ground-bdf2-b4096 chain10-bdf1-b1024 hand20-adjoint-bdf1-b2048 chain64-bdf1-b8192; do
  timeout 300 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu > $O/bench_$w.log 2>&1
  python - $O/bench_$w.log $w <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('BENCH %-28s %8.3f ms  %.3f M rollout-steps/s e2e %.3f' % (sys.argv[2], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))
PY
done
timeout 600 python tools/straggler_scaling.py > $O/straggler_scaling.log 2>&1; cat $O/straggler_scaling.log
